#!/bin/bash
# round 2, session 13: N1 projection GEMMs -- parity tests, timing, plus a regression of the fused-loss tests (ABI v4)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py -x -q > gpurun_out/s13_linear_tests.log 2>&1; echo "linear tests rc=$?"
tail -5 gpurun_out/s13_linear_tests.log
timeout 300 python tools/bench_linear.py > gpurun_out/s13_bench_linear.json 2> gpurun_out/s13_bench_linear.err; echo "bench rc=$?"
cat gpurun_out/s13_bench_linear.json | head -60
timeout 900 python -m pytest tests/test_joint_gpu.py tests/test_hardening_gpu.py -x -q > gpurun_out/s13_joint_tests.log 2>&1; echo "joint tests rc=$?"
tail -3 gpurun_out/s13_joint_tests.log
