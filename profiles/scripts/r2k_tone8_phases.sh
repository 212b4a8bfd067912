#!/bin/bash
# round 2: per-phase cycles of tone8_kernel (debug key 13, thread 0, clock64) under load
mkdir -p gpurun_out
timeout 900 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 13=1 > gpurun_out/r2k_phases.json 2> gpurun_out/r2k_phases.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2k_phases.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"]); print(json.dumps(d.get("fine_search_tier1"), indent=1))
PY
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct > gpurun_out/r2k_plain.json 2> gpurun_out/r2k_plain.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2k_plain.json
