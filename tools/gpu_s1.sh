#!/bin/bash
# round-2 GPU session 1: regression check of the hardened library + error statistics for the tight parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
timeout 600 python tools/explore_grad_error.py > gpurun_out/s1_explore.log 2>&1; echo "rc=$?" >> gpurun_out/s1_explore.log
timeout 900 python tools/explore_grad_error.py config4 > gpurun_out/s1_explore_c4.log 2>&1; echo "rc=$?" >> gpurun_out/s1_explore_c4.log
timeout 300 python tools/chunk_sweep.py -30 > gpurun_out/s1_chunk_sweep_pruned.log 2>&1
timeout 300 python tools/chunk_sweep.py 0 > gpurun_out/s1_chunk_sweep_dense.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -3 gpurun_out/s1_pytest.log; tail -5 gpurun_out/s1_explore.log; tail -12 gpurun_out/s1_explore_c4.log; cat gpurun_out/s1_chunk_sweep_pruned.log; head -c 600 gpurun_out/s1_bench.json
