"""Top source lines by warp-stall samples / executed instructions for one launch of an .ncu-rep.
usage: ncu_source_hot.py rep launch_index [topn]"""
import csv, os, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", kid,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = "?"; fn = ""; items = []; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8 or r[2] != "-": continue
    try: s = float(r[hdr.index("# Samples")]); n = float(r[hdr.index("Instructions Executed")])
    except Exception: continue
    sb = {k: r[hdr.index(k)] for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_math", "stall_mio", "stall_wait", "stall_lg") if k in hdr}
    items.append((s, n, cur, r[0], r[1].strip(), sb))
tot = sum(i[0] for i in items); toti = sum(i[1] for i in items)
print(fn[:60], "| samples", tot, "| warp instr", toti)
for s, n, f, l, src, sb in sorted(items, reverse=True)[:topn]:
    top = sorted(((float(v), k[6:]) for k, v in sb.items()), reverse=True)[:2]
    print(f"{100*s/max(tot,1):5.1f}% smp {100*n/max(toti,1):5.1f}% ins {f}:{l:>4s} {src[:90]}  [{', '.join(f'{k} {int(v)}' for v, k in top)}]")
