#!/bin/bash
# round 2, step ab: compute-sanitizer memcheck / racecheck over the tests that drive the rewritten kernels (tone8, fine_core8, TMA-ring column sums,
# grouped fallback kernels, staggered submit/collect)
mkdir -p gpurun_out
K="tone8 or osr8_fast_path_equals or trickle or tier3 or submit_collect_matches"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/r2ab_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2ab_memcheck.log | tail -3
timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tone8 or osr8_fast_path_equals or trickle" > gpurun_out/r2ab_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2ab_racecheck.log | tail -5
