"""usage: python tools/ncu_cols.py rep.ncu-rep 'Column Name' [topn] : per-source-line values of one column, sorted."""
import csv, io, subprocess, sys
rep, colname = sys.argv[1], sys.argv[2]; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
ci = {h: i for i, h in enumerate(rows[hi])}
if colname not in ci:
    print("columns:", [h for h in rows[hi]]); sys.exit(1)
c = ci[colname]; fname = ""; out = []
for r in rows[:hi] + rows[hi + 1:]:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if not r or not r[0].isdigit(): continue
    try: out.append((fname, int(r[0]), r[1], float(r[c] or 0)))
    except ValueError: pass
tot = sum(o[3] for o in out) or 1
print(f"total {colname}: {tot:.0f}")
for f, ln, src, v in sorted(out, key=lambda o: -o[3])[:topn]:
    print(f"{100*v/tot:5.1f}% {v:12.0f} {f}:{ln}: {src.strip()[:100]}")
