#!/bin/bash
# round 2, step ah: tone8_kernel capped at 48 registers (5 blocks per SM instead of 4: its two buffers are 39 KB) + the new full-size submit/collect test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tone8 or full_size or osr8 or trickle" > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ah_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct --debug 14=1 > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ah_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["stage_ms"].items()}, d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None)
PY
grep "gsmcal timeline" gpurun_out/r2ah_bench.err | tail -2
