#!/bin/bash
# round 2, step g: latency trims in fine_core8 / tone8 (parallel initial loads, fp32 band-centre atan2, one barrier less), checks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
timeout 600 python tests/stress_parity.py > gpurun_out/r2g_stress.txt 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/r2g_stress.txt
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-r-correct"
timeout 600 $B > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
timeout 600 $B --no-oracle-check --debug 13=1 > gpurun_out/r2g_bench_prof.json 2> gpurun_out/r2g_bench_prof.err; echo "bench prof rc=$?"
python - <<'PY'
import json
for f in ("r2g_bench","r2g_bench_prof"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["stage_ms"].items()}, d["oracle_agreement"]["oracle_agrees"] if d.get("oracle_agreement") else None)
    print(json.dumps(d["fine_search_tier1"].get("core8_phase_cycles_per_block")))
PY
S="python bench.py --streams 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-oracle-check --configs off --no-r-correct --pipeline 1 --groups 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2g_launches.csv $S > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"fine_core8|tone8|sch_corr|fine_peak_band|coarse_chain|colsum" -c 8 -f -o gpurun_out/prof_r2g $S > gpurun_out/r2g_ncu.log 2>&1; echo "ncu full rc=$?"
