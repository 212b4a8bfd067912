#!/bin/bash
# round 2, step ae: two GPUs, record gather asynchronous on a high-priority NCCL stream (it no longer delays the next batch's front)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct --debug 14=1 > gpurun_out/r2ae_bench_2gpu.json 2> gpurun_out/r2ae_bench_2gpu.err; echo "bench 2gpu rc=$?"
grep "gsmcal timeline" gpurun_out/r2ae_bench_2gpu.err | tail -4
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2ae_bench_2gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], d["scaling"], d["config"]["streams_per_gpu"]); print("agree", d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None)
PY
