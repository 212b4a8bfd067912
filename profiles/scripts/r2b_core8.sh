#!/bin/bash
# round 2, step b: osr-8 fast path (FIR once per burst + filtered-window cache), small-footprint burst chain.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 600 python tests/stress_parity.py > gpurun_out/r2b_stress.txt 2>&1; echo "stress rc=$?"; tail -4 gpurun_out/r2b_stress.txt
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off"
for sb in 128 512 1024; do
  timeout 600 $B --sub-batch $sb > gpurun_out/r2b_bench_sb$sb.json 2> gpurun_out/r2b_bench_sb$sb.err; echo "bench sb=$sb rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_bench_sb$sb.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","stage_ms","streams_fully_calibrated","fine_search_64bin_tier2_bursts","fine_search_allbin_fallback_bursts")}, d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None, d["synchronous_call"])
PY
done
# ncu: launch list of one synchronous step (16 streams) and --set full of the per-burst kernels
S="python bench.py --streams 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-oracle-check --configs off --pipeline 1 --groups 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_launches.csv $S > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"fine_core8|tone_est|sch_corr|fine_peak_band|coarse_chain" -c 8 -f -o gpurun_out/prof_r2b $S > gpurun_out/r2b_ncu.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_r2b.ncu-rep
