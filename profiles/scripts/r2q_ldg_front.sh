#!/bin/bash
# round 2: staggered batches, whole front at high priority; column sums by the register-path persistent kernel (no shared-memory traffic) vs the TMA ring
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
PY
  grep "gsmcal timeline" gpurun_out/r2q_$name.err | tail -3
}
run gate_hi_plain --debug 16=1
run gate_p1_t128 --debug 16=1 --persist-colsum 1 --debug 15=128
run gate_p1_t64 --debug 16=1 --persist-colsum 1 --debug 15=64
run gate_p2_t64 --debug 16=1 --persist-colsum 2 --debug 15=64
run gate_p1_t256 --debug 16=1 --persist-colsum 1 --debug 15=256
run gate_p2_t128 --debug 16=1 --persist-colsum 2 --debug 15=128
run gate_t2x3 --debug 16=1 --debug 17=2 --debug 18=3
