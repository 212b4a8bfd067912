#!/bin/bash
# round 2, step x: stream groups inside a submitted batch (debug key 8) in the staggered pipeline
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2x_$name.json 2> gpurun_out/r2x_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2x_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
PY
  grep "gsmcal timeline" gpurun_out/r2x_$name.err | tail -2
}
run g1
run g2 --debug 8=2
run g4 --debug 8=4
run g2_sb1024 --debug 8=2 --sub-batch 1024
