#!/bin/bash
# round 2: L2 prefetch for the block pf_dist ahead (fine_core8_kernel raw bytes, tone8_kernel cached window), alone and with the staggered front
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2o_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], {k: round(v, 2) for k, v in d["stage_ms"].items()})
PY
  grep "gsmcal timeline" gpurun_out/r2o_$name.err | tail -3
}
run pf592 --debug 20=592
run pf296 --debug 20=296
run pf1184 --debug 20=1184
run gate_t3x2_pf592 --debug 16=1 --debug 17=3 --debug 18=2 --debug 20=592
run gate_t2x3_pf592 --debug 16=1 --debug 17=2 --debug 18=3 --debug 20=592
run gate_t3x2_pf1184 --debug 16=1 --debug 17=3 --debug 18=2 --debug 20=1184
