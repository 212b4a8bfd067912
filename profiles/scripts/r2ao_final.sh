#!/bin/bash
# round 2, step ao: the final tree of round 2 - test suite, stress parity, the complete default bench line, the reference arm, ncu launch list and --set full
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ao_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2ao_pytest.log
tail -3 gpurun_out/r2ao_pytest.log
timeout 600 python tests/stress_parity.py > gpurun_out/r2ao_stress.txt 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/r2ao_stress.txt
timeout 1800 python bench.py > gpurun_out/r2ao_bench_full.json 2> gpurun_out/r2ao_bench_full.err; echo "bench full rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ao_bench_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "steps", d["steps"], d["stage_ms"]); print("e2e", d["e2e"]["value"], d["e2e"].get("pageable")); print("with_r", d["with_r_correct"]["value"] if d.get("with_r_correct") else None)
print(d.get("oracle_agreement", {}).get("oracle_agrees") if d.get("oracle_agreement") else None, d["clocks"], d["gpu_launches"], d.get("cpu_baseline"))
print({k:(v.get("dropin_chain_MSps"), v.get("calibrate_batch_host_MSps")) for k,v in d["configs"].items() if k in "13"}, d["configs"].get("2"), d["configs"].get("4"))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ao_bench_ref.json 2> gpurun_out/r2ao_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2ao_bench_ref.json
N="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-oracle-check --configs off --no-r-correct"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2ao_launches.csv $N > /dev/null 2>&1; echo "ncu list rc=$?"
S="python bench.py --streams 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-oracle-check --configs off --no-r-correct --pipeline 1 --groups 1"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"fine_core8|tone8|sch_corr|fine_peak_band|coarse_chain|colsum" -c 8 -f -o gpurun_out/prof_r2ao $S > gpurun_out/r2ao_ncu.log 2>&1; echo "ncu full rc=$?"
T="python bench.py --streams 64 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-oracle-check --configs off --no-r-correct --sub-batch 32"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"colsum_u8_trickle" -c 1 -f -o gpurun_out/prof_r2ao_trickle $T > gpurun_out/r2ao_ncu_trickle.log 2>&1; echo "ncu trickle rc=$?"
