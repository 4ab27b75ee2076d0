"""Summaries of ncu output for profiles/ (read here, on the CPU box).

  python tools/ncu_summary.py launches <launches.csv> <out.md>     per-kernel share of a `--metrics gpu__time_duration.sum` launch list
  python tools/ncu_summary.py full <report.ncu-rep> <out.md>      one row per profiled launch of a `--set full` capture
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_%"),
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("parq::", "")


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hi], rows[hi + 1:]
    kn, mv, gs, bs = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        a = agg.setdefault(short(r[kn]), [0, 0.0, r[gs], r[bs]])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total us | mean us | share | last grid | block |\n|---|---|---|---|---|---|---|\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.1f | %.2f | %.1f%% | %s | %s |\n" % (n[:70], a[0], a[1] / 1e3, a[1] / 1e3 / a[0], 100 * a[1] / tot, a[2], a[3]))
        f.write("\n%d launches, %.1f us total (cold-cache, serialised under ncu: compare shares, not absolutes)\n" % (len(data), tot / 1e3))


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    H, U = rows[0], rows[1]
    kn = H.index("Kernel Name")
    cols = [(H.index(m), lab) for m, lab in METRICS if m in H]
    with open(out, "w") as f:
        f.write("| kernel | " + " | ".join("%s [%s]" % (lab, U[i]) if U[i] else lab for i, lab in cols) + " |\n")
        f.write("|---|" + "---|" * len(cols) + "\n")
        for r in rows[2:]:
            if len(r) != len(H):
                continue
            f.write("| `%s` | " % short(r[kn])[:60] + " | ".join(r[i] for i, _ in cols) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
