#!/bin/bash
# session 28: persistent warp-specialised projection kernel -- parity suites, canaries, timing against the simple kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_linear_gpu.py -x -q > gpurun_out/s28_linear_tests.log 2>&1; echo "linear tests rc=$?"; tail -4 gpurun_out/s28_linear_tests.log
timeout 300 python -m pytest tests/test_hardening_gpu.py tests/test_predictor_gpu.py -q -k "projection or predictor" > gpurun_out/s28_more_tests.log 2>&1; echo "canary/predictor tests rc=$?"; tail -3 gpurun_out/s28_more_tests.log
echo "== persistent"; timeout 200 python tools/bench_linear.py 2>/dev/null | grep -E "us_ours|linear_gemm|fold" 
echo "== simple"; TSASR_LINEAR_SIMPLE=1 timeout 200 python tools/bench_linear.py 2>/dev/null | grep -E "us_ours|linear_gemm|fold"
