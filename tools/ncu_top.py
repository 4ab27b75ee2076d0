"""Top stall instructions of a kernel from an ncu report (source page).  Usage:
   python tools/ncu_top.py <report.ncu-rep> <kernel regex> [launch index] [top N]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = rows[1]
data = [r for r in rows[2:] if len(r) == len(H) and r[0] != "Address"]
si, src, ex = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data)
print(rows[0][1][:100], "total samples", tot, "instructions", len(data))
for r in sorted(data, key=lambda r: -int(r[si]))[:topn]:
    st = sorted([(int(r[i]), H[i][6:]) for i in stall_cols], reverse=True)[:2]
    print("%6d %5.1f%% ex=%9s %-64s %s" % (int(r[si]), 100 * int(r[si]) / max(tot, 1), r[ex], r[src].strip()[:64], st))
