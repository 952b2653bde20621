#!/bin/bash
# session 15: N3 predictor kernels (first run: bounded by timeouts), N1 tests after the conversion change, timings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_predictor_gpu.py -x -q > gpurun_out/s15_predictor_tests.log 2>&1; echo "predictor tests rc=$?"
tail -25 gpurun_out/s15_predictor_tests.log
timeout 300 python -m pytest tests/test_linear_gpu.py -q > gpurun_out/s15_linear_tests.log 2>&1; echo "linear tests rc=$?"
tail -4 gpurun_out/s15_linear_tests.log
timeout 200 python tools/bench_predictor.py > gpurun_out/s15_bench_predictor.json 2> gpurun_out/s15_bench_predictor.err; echo "bench predictor rc=$?"
cat gpurun_out/s15_bench_predictor.json; tail -3 gpurun_out/s15_bench_predictor.err
timeout 200 python tools/bench_linear.py > gpurun_out/s15_bench_linear.json 2> gpurun_out/s15_bench_linear.err; echo "bench linear rc=$?"
grep -E "us_|kernel" gpurun_out/s15_bench_linear.json | head -40
