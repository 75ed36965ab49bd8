"""Static SASS mnemonic counts per kernel of libpixo_b200.so (cuobjdump -sass) -> profiles/*.txt.
usage: python tools/sass_table.py [out.txt]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("PIXO_B200_SO") or os.path.join(ROOT, "pixo_b200", "libpixo_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt")
COLS = ["UTMALDG", "SYNCS", "FADD2", "FMUL2", "FFMA2", "IDP", "VABSDIFF4", "VIMNMX", "REDUX", "LDGSTS", "ATOMS", "ATOMG", "RED",
        "SHFL", "PRMT", "BMSK", "FLO", "STL", "LDL"]
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, tab = None, collections.OrderedDict()
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        mangled = m.group(1)
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", dem.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("void ", "").replace("pixo::", "")
        name = short
        tab[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        tab[name]["total"] += 1
        tab[name][op] += 1
with open(out, "w") as f:
    f.write("# static SASS mnemonic counts per kernel of libpixo_b200.so (cuobjdump -sass, sm_100a); UTMALDG = TMA tensor load,\n"
            "# SYNCS = mbarrier ops, FADD2/FMUL2/FFMA2 = packed f32x2, IDP = dp4a, VABSDIFF4 = byte SAD, REDUX = warp reduce,\n"
            "# LDGSTS = cp.async, ATOMS/ATOMG/RED = shared / global atomics, BMSK/FLO = bit-mask / find-leading-one, STL/LDL = local memory\n")
    f.write(f"{'kernel':34s}" + "".join(f"{c:>10s}" for c in ["total"] + COLS) + "\n")
    for k, c in sorted(tab.items(), key=lambda kv: kv[0]):
        f.write(f"{k[:34]:34s}" + "".join(f"{c[x]:10d}" for x in ["total"] + COLS) + "\n")
print(open(out).read())
