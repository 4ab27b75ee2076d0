#!/bin/bash
# Multi-GPU visit: bench line and config-3 eval pipeline on N GPUs of one box.   usage: bash tools/gpu_multi.sh <tag> <N> <clips> [steps]
TAG=$1; N=$2; CLIPS=$3; STEPS=${4:-20}
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 5 \
  > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
tail -c 800 gpurun_out/bench_${TAG}_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_${N}gpu.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "pipeline", d["e2e_pipeline"]["value"]); print(d["detection_gather"])
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/eval_pipeline.py --clips $CLIPS \
  > gpurun_out/eval_${TAG}_${N}gpu.json 2> gpurun_out/eval_${TAG}_${N}gpu.err
cat gpurun_out/eval_${TAG}_${N}gpu.json; tail -c 500 gpurun_out/eval_${TAG}_${N}gpu.err
echo multi-done
