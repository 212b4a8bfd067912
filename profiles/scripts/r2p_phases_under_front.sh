#!/bin/bash
# round 2: phase cycles of fine_core8_kernel / tone8_kernel for a batch that ran under the next batch's front (trickle column sums + burst chain)
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1 --debug 13=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
pp=d["pipelined_phase_cycles_per_block"]
for k in ("core8","tone8_fine","tone8_post"):
    a,b=pp["batch_before_last"][k],pp["last_batch"][k]
    print(k); [print("   %-20s %9.1f %9.1f  %+6.1f%%" % (n,a[n],b[n],100*(a[n]-b[n])/max(b[n],1))) for n in a]
PY
  grep "gsmcal timeline" gpurun_out/r2p_$name.err | tail -2
}
run gate_t2x3 --debug 16=1 --debug 17=2 --debug 18=3
run gate_t3x2 --debug 16=1 --debug 17=3 --debug 18=2
run gate_t7x1 --debug 16=1 --debug 17=7 --debug 18=1
