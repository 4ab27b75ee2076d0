#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list and full captures of the hot kernels.
# usage: tools/gpu_round.sh <tag> [stages: test bench launches full]
TAG=${1:-rX}; shift
STAGES=${@:-test bench launches full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt
for s in $STAGES; do
  case $s in
    test)
      timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/test_$TAG.log ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
      tail -c 600 gpurun_out/bench_$TAG.err; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "attn frac", d["roofline"]["frac"], "samp frac", d["roofline_sampling"]["frac"])
print(d["breakdown_ms_per_step"]); print(d["clocks"]); print(d.get("cpu_baseline"))
PY
      ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
        python bench.py --steps 3 --warmup 3 > gpurun_out/launches_$TAG.log 2>&1
      tail -3 gpurun_out/launches_$TAG.log ;;
    full)
      timeout 900 ncu --set full --clock-control none --import-source on \
        -k regex:"attn[23]?_tc|attn3_combine|project_sample|attn_combine|heads_final|add_ln|gn_apply|posemb" -c 14 -f -o gpurun_out/prof_${TAG}_iter \
        python tools/prof_step.py 1 > gpurun_out/prof_${TAG}_iter.log 2>&1
      tail -2 gpurun_out/prof_${TAG}_iter.log
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm2?_tc" -c 6 -f -o gpurun_out/prof_${TAG}_gemm \
        python tools/prof_step.py 1 > gpurun_out/prof_${TAG}_gemm.log 2>&1
      tail -2 gpurun_out/prof_${TAG}_gemm.log ;;
  esac
done
