#!/bin/bash
# round 2, step ap: generic loader - the derotation step phasor once per block instead of once per thread
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ap_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ap_pytest.log
timeout 600 python tests/stress_parity.py > gpurun_out/r2ap_stress.txt 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/r2ap_stress.txt
for n in 1024; do
timeout 900 python bench.py --streams $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct --debug 14=1 > gpurun_out/r2ap_bench_$n.json 2> gpurun_out/r2ap_bench_$n.err; echo "bench $n rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2ap_bench_$n.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["stage_ms"].items()}, d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None)
PY
grep "gsmcal timeline" gpurun_out/r2ap_bench_$n.err | tail -2
done
