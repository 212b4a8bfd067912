#!/bin/bash
# round 2, step u: per-step phase cycles of coarse_chain_kernel (debug key 13)
mkdir -p gpurun_out
for mode in "default" "lockstep --debug 16=0 --debug 17=0"; do
set -- $mode; name=$1; shift
timeout 600 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 13=1 "$@" > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err; echo "== $name rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2u_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["stage_ms"])
pp=d["pipelined_phase_cycles_per_block"]
for k in pp: print(k, pp[k]["coarse_chain_per_step"])
PY
done
