"""Top source lines / SASS instructions by warp-stall samples for one launch of an .ncu-rep.
usage: ncu_source_hot.py rep launch_index [topn] [cuda|sass]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
view = sys.argv[4] if len(sys.argv) > 4 else "cuda"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view, "--launch-skip", kid,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and "# Samples" in r:
        hdr = r; start = i + 1; break
if hdr is None:
    print(out[:1500]); sys.exit(1)
si, sa, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
first = hdr[0]
tot = 0; items = []
for r in rows[start:]:
    if len(r) <= sa: continue
    try: v = float(r[sa])
    except Exception: continue
    tot += v
    items.append((v, r[0], r[si].strip(), r[ie]))
items.sort(reverse=True)
print(rows[0][1] if len(rows[0]) > 1 else "", "| total samples", tot)
for v, l, s, n in items[:topn]:
    print(f"{v:7.0f} {100*v/max(tot,1):5.1f}% inst={n:>9s} {first}={l[-6:]:>6s} {s[:120]}")
