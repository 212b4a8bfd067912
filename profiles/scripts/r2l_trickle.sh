#!/bin/bash
# round 2: column sums of batch k+1 by a one-warp-per-SM TMA-ring kernel under the FP64 stages of batch k (staggered batches)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "submit" 2>&1 | tail -5
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2l_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
PY
  grep "gsmcal timeline" gpurun_out/r2l_$name.err | tail -4
}
run default
run gate_t7 --debug 16=1 --debug 17=7
run gate_t7_sch56 --debug 16=1 --debug 17=7 --debug 19=1
run nogate_t7_sch56 --debug 17=7 --debug 19=1
run gate_t3x2_sch56 --debug 16=1 --debug 17=3 --debug 18=2 --debug 19=1
run gate_t4_sch56 --debug 16=1 --debug 17=4 --debug 19=1
run gate_t7_sch56_d3 --debug 16=1 --debug 17=7 --debug 19=1 --pipeline 3
run gate_t7_sch56_sb256 --debug 16=1 --debug 17=7 --debug 19=1 --sub-batch 256
