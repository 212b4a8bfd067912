#!/bin/bash
# round 2: staggered batches + trickle column sums, per-stage device timeline
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["clocks"])
PY
  grep "gsmcal timeline" gpurun_out/r2n_$name.err | tail -4
}
run default
run gate_only --debug 16=1
run gate_t3x2_sch56 --debug 16=1 --debug 17=3 --debug 18=2 --debug 19=1
run gate_t3x2_sch64 --debug 16=1 --debug 17=3 --debug 18=2
run gate_t2x3_sch56 --debug 16=1 --debug 17=2 --debug 18=3 --debug 19=1
run gate_t3x2_sch56_sb1024 --debug 16=1 --debug 17=3 --debug 18=2 --debug 19=1 --sub-batch 1024
run gate_t3x2_sch56_d3 --debug 16=1 --debug 17=3 --debug 18=2 --debug 19=1 --pipeline 3
