#!/bin/bash
# round 2: device timeline of the submit/collect pipeline - does the front of batch k+1 run under the FP64 stages of batch k?
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
PY
  grep "gsmcal timeline" gpurun_out/r2i_$name.err | tail -8
}
run default
run p1_256 --persist-colsum 1
run p1_128 --persist-colsum 1 --debug 15=128
run p2_128 --persist-colsum 2 --debug 15=128
run p4_64 --persist-colsum 4 --debug 15=64
