"""Per-kernel SASS census of libparq_b200.so: tcgen05 / TMEM / TMA instruction counts from `cuobjdump -sass`
(B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk).
    python tools/sass_census.py > profiles/r2_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parq_b200", "libparq_b200.so")
PATTERNS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
            ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("HMMA/IMMA (legacy mma.sync)", r"\b[HI]MMA\b"), ("LDL/STL (local memory)", r"\b(LDL|STL)\b")]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", line):
            kernels[cur]["instructions"] += 1
            for name, pat in PATTERNS:
                if re.search(pat, line):
                    kernels[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    print("SASS census of parq_b200/libparq_b200.so (sm_100a), `cuobjdump -sass`, one row per kernel\n")
    print("| kernel | SASS instructions | " + " | ".join(n for n, _ in PATTERNS) + " |")
    print("|---|---|" + "---|" * len(PATTERNS))
    for (k, c), d in zip(kernels.items(), demangle):
        d = re.sub(r"\(.*", "", d).replace("void ", "").replace("parq::", "")
        print("| `%s` | %d | " % (d[:70], c["instructions"]) + " | ".join(str(c[n]) if c[n] else "" for n, _ in PATTERNS) + " |")


if __name__ == "__main__":
    main()
