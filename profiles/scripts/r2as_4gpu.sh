#!/bin/bash
# round 2, step as: four GPUs of one box, device-resident leg only (256 streams per GPU: one batch per step)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --configs off --no-r-correct --no-oracle-check > gpurun_out/r2as_bench_4gpu.json 2> gpurun_out/r2as_bench_4gpu.err; echo "bench 8gpu rc=$?"
tail -2 gpurun_out/r2as_bench_4gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2as_bench_4gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], d["scaling"], d["config"]["streams_per_gpu"], d["n_gpus"], d["clocks"])
PY
