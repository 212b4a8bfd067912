#!/bin/bash
# round 2, step d: 56-register FP64 kernels (room for a column-sum block beside them), slide only where the maximum can be, 8 passes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
timeout 600 python tests/stress_parity.py > gpurun_out/r2d_stress.txt 2>&1; echo "stress rc=$?"; tail -4 gpurun_out/r2d_stress.txt
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2d_bench_$name.json 2> gpurun_out/r2d_bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2d_bench_$name.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","stage_ms","streams_fully_calibrated","fine_search_64bin_tier2_bursts","fine_search_allbin_fallback_bursts","fine_search_tier1")}, d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None, d.get("synchronous_call"))
except Exception as e: print("parse failed", e)
PY
}
run sb512 --sub-batch 512
run sb256 --sub-batch 256 --no-oracle-check
run sb128 --sub-batch 128 --no-oracle-check
run sb512_persist1 --sub-batch 512 --persist-colsum 1 --no-oracle-check
run sb256_persist1 --sub-batch 256 --persist-colsum 1 --no-oracle-check
run sb256_depth3 --sub-batch 256 --pipeline 3 --no-oracle-check
S="python bench.py --streams 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-oracle-check --configs off --pipeline 1 --groups 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2d_launches.csv $S > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"fine_core8|tone8|sch_corr" -c 4 -f -o gpurun_out/prof_r2d $S > gpurun_out/r2d_ncu.log 2>&1; echo "ncu full rc=$?"
