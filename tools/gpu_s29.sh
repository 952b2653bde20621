#!/bin/bash
# session 29: 16-producer form of the joint kernel for narrow vocabularies (V <= 128) -- parity suites, A/B timing at the recipe's shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_joint_gpu.py tests/test_hardening_gpu.py -x -q > gpurun_out/s29_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s29_tests.log
{
for shape in "16 400 240 640 29" "16 400 100 640 100" "16 400 100 640 1000"; do
  echo "== shape $shape"
  echo "-- narrow form (default)"; timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
  echo "-- TSASR_DEBUG_NO_NARROW=1"; TSASR_DEBUG_NO_NARROW=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
done
} > gpurun_out/s29_ab_narrow.txt 2>&1
cat gpurun_out/s29_ab_narrow.txt
