#!/bin/bash
# round 2, last check of the final tree: GPU suite, smoke(), stress parity, short default bench with the oracle-agreement leg
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ar_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ar_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python tests/stress_parity.py > gpurun_out/r2ar_stress.txt 2>&1; echo "stress rc=$?"; tail -1 gpurun_out/r2ar_stress.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct > gpurun_out/r2ar_bench.json 2> gpurun_out/r2ar_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ar_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["stage_ms"].items()}, d["oracle_agreement"]["oracle_agrees"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
