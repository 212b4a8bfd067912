#!/bin/bash
# round 2: staggered batches (FP64 stages of batch k+1 gated behind batch k) x column-sum variants; device timeline on stderr
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2j_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
PY
  grep "gsmcal timeline" gpurun_out/r2j_$name.err | tail -4
}
run gate_default
run gate_p1_256 --persist-colsum 1
run gate_p2_128 --persist-colsum 2 --debug 15=128
run gate_p2_256 --persist-colsum 2
run gate_p1_256_sb256 --persist-colsum 1 --sub-batch 256
run gate_p1_256_sb256_d3 --persist-colsum 1 --sub-batch 256 --pipeline 3
run nogate_default --debug 16=0
