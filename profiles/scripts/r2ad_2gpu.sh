#!/bin/bash
# round 2, step ad (final tree): two GPUs of one box - strong-scaling bench line under torchrun, reference arm rank handling, multi-device test
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2ad_topo_2gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "second_device or submit_cancel or pageable or bench_workload" > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ad_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ad_bench_2gpu.json 2> gpurun_out/r2ad_bench_2gpu.err; echo "bench 2gpu rc=$?"
tail -3 gpurun_out/r2ad_bench_2gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2ad_bench_2gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], d["scaling"], d["config"]["streams_per_gpu"], d["stage_ms"]); print("e2e", json.dumps(d["e2e"])[:900]); print("with_r", d["with_r_correct"]); print("agree", d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None)
PY
