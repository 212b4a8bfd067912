#!/bin/bash
# round 2, step aa: the per-GPU share of config 5 at N = 8, 4, 2 GPUs (128 / 256 / 512 streams) on one GPU: does the staggered pipeline
# stay FP64-bound when a step is a single small batch?
mkdir -p gpurun_out
for n in 128 256 512; do
for mode in "default" "lockstep --debug 16=0 --debug 17=0"; do
set -- $mode; name=$1; shift
timeout 600 python bench.py --streams $n --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1 "$@" > gpurun_out/r2aa_${n}_$name.json 2> gpurun_out/r2aa_${n}_$name.err; echo "== $n streams, $name rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2aa_${n}_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], {k: round(v, 2) for k, v in d["stage_ms"].items()})
PY
grep "gsmcal timeline" gpurun_out/r2aa_${n}_$name.err | tail -2
done; done
