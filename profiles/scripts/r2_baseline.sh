#!/bin/bash
# round-2 baseline on the B200: GPU test suite, the randomised + planted-signal parity sweep, and the new config-5 bench line
# (1024 streams on one GPU) with the round-1 kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
free -g > gpurun_out/r2_mem.txt; nproc >> gpurun_out/r2_mem.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_base_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_base_pytest.log
timeout 900 python tests/stress_parity.py > gpurun_out/r2_stress_parity.txt 2>&1; echo "stress rc=$?"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_base_bench.json
tail -5 gpurun_out/r2_base_pytest.log; tail -8 gpurun_out/r2_stress_parity.txt
