#!/bin/bash
# round 2, step e: carve-out for the front kernels (column sums beside the FP64 kernels), staging ring for pageable buffers, new tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --configs off --no-oracle-check --no-r-correct"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2e_bench_$name.json 2> gpurun_out/r2e_bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_bench_$name.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, {k:round(v,2) for k,v in d["stage_ms"].items()}, d.get("synchronous_call",{}).get("ms_per_step"))
except Exception as e: print("parse failed", e)
PY
}
run sb512 --sub-batch 512
run sb512_p1 --sub-batch 512 --persist-colsum 1
run sb512_p2 --sub-batch 512 --persist-colsum 2
run sb256_p1 --sub-batch 256 --persist-colsum 1
run sb256_p1_d3 --sub-batch 256 --persist-colsum 1 --pipeline 3
run sb128_p1_d4 --sub-batch 128 --persist-colsum 1 --pipeline 4
# the complete default line (e2e pinned / bare / pageable, configs 1-4, oracle agreement)
timeout 1500 python bench.py --steps 10 --warmup 3 --sub-batch 512 > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err; echo "bench full rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2e_bench_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"]); print("e2e", json.dumps(d["e2e"])[:1500]); print("configs", json.dumps(d["configs"])[:3000]); print("agreement", d["oracle_agreement"]["oracle_agrees"], d["oracle_agreement"]["outcome_histogram_rank0"])
print("roofline", json.dumps(d["roofline"])[:800])
PY
