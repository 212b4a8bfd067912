#!/bin/bash
# round 2, step ak: what does the burst chain cost the previous batch's burst kernels - the SM slots it holds or the instructions it runs?
# An extra kernel of sleeping blocks (64 threads, N KB of shared memory, ~2.3 ms) is launched behind the chain on its high-priority stream.
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2ak_$name.json 2> gpurun_out/r2ak_$name.err; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2ak_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
PY
  grep "gsmcal timeline" gpurun_out/r2ak_$name.err | tail -2
}
run none
run occupy_1kb --debug 23=1
run occupy_24kb --debug 23=24
run occupy_48kb --debug 23=48
