"""Top sampled SASS instructions of the first kernel in an .ncu-rep matching a name substring.
usage: ncu_hot.py file.ncu-rep substr [topN]"""
import csv, io, subprocess, sys
path, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
for blk in txt.split('"Kernel Name",')[1:]:
    lines = blk.split("\n")
    if sub not in lines[0]:
        continue
    rows = list(csv.reader(lines[1:]))
    h = rows[0]
    data = [r for r in rows[1:] if len(r) == len(h)]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp]) for r in data)
    print(lines[0][:80], "samples", tot, "instrs", len(data), "executed", sum(int(r[iex]) for r in data))
    agg = {}
    for r in data:
        for c in stall:
            agg[h[c][6:]] = agg.get(h[c][6:], 0) + int(r[c])
    print("  stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
    for i in sorted(idx):
        r = data[i]
        st = {h[c][6:]: int(r[c]) for c in stall if int(r[c]) > 0}
        print(i, r[isamp], r[iex], r[isrc].strip()[:75], sorted(st.items(), key=lambda kv: -kv[1])[:2])
    break
