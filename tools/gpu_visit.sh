#!/bin/bash
# One GPU-box visit of round 2: parity tests, sampling micro-benchmark, both bench arms, one ncu capture.
# usage (on the box): bash tools/gpu_visit.sh <tag> [stages...]   stages: proj test sample ref bench ncu_sample ncu_iter ncu_gemm ncu_one_clip launches trace configs ab sanit (one ncu capture per call: gpurun_out/ is capped at 64 MiB)
TAG=${1:-r2x}; shift
STAGES=${@:-proj test sample ref bench}
mkdir -p gpurun_out
for s in $STAGES; do
  case $s in
    proj) timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "projection" 2>&1 | tail -15 > gpurun_out/test_${TAG}_proj.log; tail -6 gpurun_out/test_${TAG}_proj.log ;;
    test) timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -60 > gpurun_out/test_${TAG}.log; tail -40 gpurun_out/test_${TAG}.log ;;
    sample) timeout 300 python tools/bench_sample.py > gpurun_out/bench_sample_${TAG}.log 2>&1; tail gpurun_out/bench_sample_${TAG}.log ;;
    ref) timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; tail -c 300 gpurun_out/bench_ref_${TAG}.err ;;
    bench) timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/bench_${TAG}.err
      python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "attn frac", d["roofline"]["frac"])
    print("sampling", {k: d["roofline_sampling"][k] for k in ("frac", "ms_per_launch", "in_step")})
    print(d["breakdown_ms_per_step"]); print(d["breakdown_launches"]); print(d["clocks"]); print(d.get("cpu_baseline")); print(d.get("gpu_torch_baseline"))
except Exception as e:
    print("bench line unreadable:", e)
PY
      ;;
    ncu_sample) timeout 300 ncu --set full --clock-control none --import-source on -k regex:"project_sample" -c 3 -f -o gpurun_out/prof_${TAG}_sample python tools/prof_step.py 1 > gpurun_out/prof_${TAG}_sample.log 2>&1; tail -2 gpurun_out/prof_${TAG}_sample.log ;;
    ncu_iter) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn[23]?_tc|attn3_combine|project_sample|heads_final|add_ln|gn_apply|posemb|chain" -c 16 -f -o gpurun_out/prof_${TAG}_iter python tools/prof_step.py 1 > gpurun_out/prof_${TAG}_iter.log 2>&1; tail -2 gpurun_out/prof_${TAG}_iter.log ;;
    ncu_gemm) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm2?_tc" -c 8 -f -o gpurun_out/prof_${TAG}_gemm python tools/prof_step.py 1 > gpurun_out/prof_${TAG}_gemm.log 2>&1; tail -2 gpurun_out/prof_${TAG}_gemm.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 > gpurun_out/launches_${TAG}.log 2>&1; tail -3 gpurun_out/launches_${TAG}.log ;;
    ncu_one_clip) timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:"gemm_sk|heads_final_small|attn3_combine|attn3_tc|attn2_tc|project_sample" -s 8 -c 14 -f -o gpurun_out/prof_${TAG}_one_clip python tools/prof_step.py 2 1 > gpurun_out/prof_${TAG}_one_clip.log 2>&1; tail -2 gpurun_out/prof_${TAG}_one_clip.log ;;
    trace) timeout 200 python tools/launch_trace.py > gpurun_out/trace_${TAG}.txt 2>&1; head -16 gpurun_out/trace_${TAG}.txt
      timeout 100 python tools/launch_trace.py 1 20 > gpurun_out/trace_${TAG}_one_clip.txt 2>&1; head -2 gpurun_out/trace_${TAG}_one_clip.txt | cut -c1-300 ;;
    configs) timeout 300 python tools/bench_configs.py > gpurun_out/configs_${TAG}.json 2> gpurun_out/configs_${TAG}.err; cut -c1-400 gpurun_out/configs_${TAG}.json ;;
    ab) timeout 300 python tools/ab_step.py --err --rounds 12 base=0 qkv=7 all=0x7ff merge=0,fused_merge=1 > gpurun_out/ab_${TAG}.txt 2>&1; tail -12 gpurun_out/ab_${TAG}.txt ;;
    sanit) timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/sanitize_memcheck_${TAG}.log 2>&1; tail -2 gpurun_out/sanitize_memcheck_${TAG}.log
      timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_step.py > gpurun_out/sanitize_synccheck_${TAG}.log 2>&1; tail -2 gpurun_out/sanitize_synccheck_${TAG}.log ;;
  esac
done
echo visit-done
