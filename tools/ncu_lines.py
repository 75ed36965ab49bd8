"""Aggregate an .ncu-rep's source page by CUDA source line: samples, instructions, smem conflicts.
usage: python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
S, N, X = ci["# Samples"], ci["Instructions Executed"], ci["L1 Wavefronts Shared Excessive"]
fname = ""
out = []
for r in rows[:hi] + rows[hi + 1:]:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if not r or not r[0].isdigit(): continue
    try:
        out.append((fname, int(r[0]), r[1], int(r[S] or 0), int(r[N] or 0), int(r[X] or 0)))
    except ValueError:
        pass
ts = sum(o[3] for o in out) or 1; ti = sum(o[4] for o in out) or 1
print(f"total samples {ts} instr {ti} excessive smem wavefronts {sum(o[5] for o in out)}")
for f, ln, src, s, n, ex in sorted(out, key=lambda o: -o[3])[:topn]:
    print(f"{100*s/ts:5.1f}% smp {100*n/ti:5.1f}% ins xwf {ex:9d} {f}:{ln}: {src.strip()[:100]}")
if len(sys.argv) > 3:
    # region sums: "name:lo-hi,name:lo-hi" for the file given first "file=..."
    spec = sys.argv[3]
    for part in spec.split(","):
        name, rng = part.split(":"); lo, hi = map(int, rng.split("-"))
        s = sum(o[3] for o in out if lo <= o[1] <= hi and o[0].endswith(sys.argv[4]))
        n = sum(o[4] for o in out if lo <= o[1] <= hi and o[0].endswith(sys.argv[4]))
        print(f"region {name:12s} {100*s/ts:5.1f}% samples {100*n/ti:5.1f}% instr")
