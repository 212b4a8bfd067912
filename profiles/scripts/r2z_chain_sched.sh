#!/bin/bash
# round 2, step z: how the burst chains of batch k+1 share the SMs with the FP64 stages of batch k: register cap x serialised stream groups x priority
mkdir -p gpurun_out
C=multi-rtl-sdr-calibration_b200/csrc
cp $C/libgsmcal.so /tmp/libgsmcal_keep.so
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 14=1"
run() { r=$1; name=$2; shift; shift
cp $C/libgsmcal_chain$r.so $C/libgsmcal.so
timeout 600 $B "$@" > gpurun_out/r2z_$name.json 2> gpurun_out/r2z_$name.err; echo "== $name rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2z_$name.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])
PY
grep "gsmcal timeline" gpurun_out/r2z_$name.err | tail -2
}
run 64 c64_serial --debug 21=1
run 64 c64_serial_lo --debug 21=1 --debug 22=1
run 64 c64_lo --debug 22=1
run 64 c64_serial_g4 --debug 21=1 --debug 8=4
run 112 c112_serial --debug 21=1
run 112 c112_lo --debug 22=1
cp /tmp/libgsmcal_keep.so $C/libgsmcal.so
