#!/bin/bash
# session 36: dW loader skips the dY images beyond the vocabulary -- backward parity suites, timing at the recipe's shape and at config 2 / config 4 widths
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_joint_gpu.py tests/test_hardening_gpu.py tests/test_parity_tight_gpu.py -x -q > gpurun_out/s36_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s36_tests.log
{
for shape in "16 400 240 640 29" "16 400 100 640 1000" "4 400 100 640 5000"; do
  echo "== shape $shape"; timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
done
} > gpurun_out/s36_dw_skip.txt 2>&1
cat gpurun_out/s36_dw_skip.txt
