#!/bin/bash
# round 2, step y: register cap of coarse_chain_kernel (how many chains fit into what one retiring FP64 block frees); 2 stream groups per batch
mkdir -p gpurun_out
C=multi-rtl-sdr-calibration_b200/csrc
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
for r in 112 88 64; do
cp $C/libgsmcal_chain$r.so $C/libgsmcal.so
timeout 600 $B > gpurun_out/r2y_chain$r.json 2> gpurun_out/r2y_chain$r.err; echo "== chain $r regs rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_chain$r.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], {k: round(v, 2) for k, v in d["stage_ms"].items()})
PY
grep "gsmcal timeline" gpurun_out/r2y_chain$r.err | tail -2
done
