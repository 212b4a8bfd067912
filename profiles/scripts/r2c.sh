#!/bin/bash
# round 2, step c: multi-pass tier 1, cache-fed tone8 kernel, SCH correlation at 4 blocks/SM, Appendix-A + MEX GPU tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
timeout 600 python tests/stress_parity.py > gpurun_out/r2c_stress.txt 2>&1; echo "stress rc=$?"; tail -4 gpurun_out/r2c_stress.txt
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --sub-batch 512"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2c_bench_$name.json 2> gpurun_out/r2c_bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c_bench_$name.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","stage_ms","streams_fully_calibrated","fine_search_64bin_tier2_bursts","fine_search_allbin_fallback_bursts","fine_search_tier1")}, d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None)
except Exception as e: print("parse failed", e)
PY
}
run default
run passes2 --debug 10=2 --no-oracle-check
run passes8 --debug 10=8 --no-oracle-check
run notone8 --debug 11=1 --no-oracle-check
S="python bench.py --streams 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-oracle-check --configs off --pipeline 1 --groups 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches.csv $S > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"fine_core8|tone8|tone_est|sch_corr|fine_peak_band" -c 9 -f -o gpurun_out/prof_r2c $S > gpurun_out/r2c_ncu.log 2>&1; echo "ncu full rc=$?"
