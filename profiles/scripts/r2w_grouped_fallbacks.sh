#!/bin/bash
# round 2, step w: fallback kernels (tone_est, tier 2) own groups of 8 bursts per block behind their need-masks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2w_pytest.log
timeout 600 python tests/stress_parity.py > gpurun_out/r2w_stress.txt 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/r2w_stress.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct --debug 14=1 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2w_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["stage_ms"].items()}, d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None, d["gpu_launches"])
PY
grep "gsmcal timeline" gpurun_out/r2w_bench.err | tail -2
S="python bench.py --streams 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-oracle-check --configs off --no-r-correct --pipeline 1 --groups 1"
