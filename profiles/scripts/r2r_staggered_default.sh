#!/bin/bash
# round 2, step r: staggered batches + TMA-ring column sums as the submit/collect default - full checks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2r_pytest.log
tail -4 gpurun_out/r2r_pytest.log
timeout 600 python tests/stress_parity.py > gpurun_out/r2r_stress.txt 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/r2r_stress.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct --debug 14=1 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["stage_ms"].items()}, d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None, d["gpu_launches"])
PY
grep "gsmcal timeline" gpurun_out/r2r_bench.err | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct --no-oracle-check --debug 16=0 --debug 17=0 > gpurun_out/r2r_bench_lockstep.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2r_bench_lockstep.json').read().strip().splitlines()[-1]); print('lockstep', d['value'], d['ms_per_step'])"
