#!/bin/bash
# round 2, step aq: the TMA-ring column sums as one launch per stream group, so the chain of group g starts under the sums of group g+1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "submit or trickle or staggered" > gpurun_out/r2aq_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2aq_pytest.log
for n in 128 256 1024; do
timeout 600 python bench.py --streams $n --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1 > gpurun_out/r2aq_$n.json 2> gpurun_out/r2aq_$n.err; echo "== $n streams rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2aq_$n.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
PY
grep "gsmcal timeline" gpurun_out/r2aq_$n.err | tail -2
done
