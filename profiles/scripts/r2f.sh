#!/bin/bash
# round 2, step f: test suite, the complete default line, fine_core8 phase cycles
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct --debug 13=1 > gpurun_out/r2f_bench_prof.json 2> gpurun_out/r2f_bench_prof.err; echo "bench prof rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f_bench_prof.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stage_ms"]); print(json.dumps(d["fine_search_tier1"]))
PY
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_full.json 2> gpurun_out/r2f_bench_full.err; echo "bench full rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f_bench_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], d["stage_ms"]); print("e2e", d["e2e"]["value"], d["e2e"].get("pageable")); print("with_r", d["with_r_correct"]["value"] if d.get("with_r_correct") else None)
print({k:(v.get("dropin_chain_MSps"), v.get("calibrate_batch_host_MSps")) for k,v in d["configs"].items() if k in "13"}, d["configs"].get("4"))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2f_bench_ref.json
