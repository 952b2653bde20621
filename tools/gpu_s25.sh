#!/bin/bash
# session 25: fuzz tests of the N1 / N3 kernels + their parity suites after the registry change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_predictor_gpu.py -q > gpurun_out/s25_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/s25_tests.log
