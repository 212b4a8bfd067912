#!/bin/bash
# round 2, step al: outputs per thread of the generic loader's FIR (LW_R 5 / 7 / 9): SCH correlation stage
mkdir -p gpurun_out
C=multi-rtl-sdr-calibration_b200/csrc
cp $C/libgsmcal.so /tmp/libgsmcal_keep.so
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct"
for r in 5 7 9 3; do
cp $C/libgsmcal_lw$r.so $C/libgsmcal.so
timeout 600 $B > gpurun_out/r2al_lw$r.json 2> gpurun_out/r2al_lw$r.err; echo "== LW_R $r rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2al_lw$r.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], {k: round(v, 2) for k, v in d["stage_ms"].items()}, d["oracle_agreement"]["oracle_agrees"])
PY
done
cp /tmp/libgsmcal_keep.so $C/libgsmcal.so
