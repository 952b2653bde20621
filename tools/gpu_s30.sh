#!/bin/bash
# session 30: narrow vocabularies -- resident W slices (no W stream) + 16 producer warps; parity suites, A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_joint_gpu.py tests/test_hardening_gpu.py -x -q > gpurun_out/s30_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s30_tests.log
{
for shape in "16 400 240 640 29" "16 400 100 640 100" "16 400 100 256 64"; do
  echo "== shape $shape"
  echo "-- default (resident W, 16 producer warps)"; timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
  echo "-- TSASR_DEBUG_NARROW_8=1 (resident W, 8 producer warps)"; TSASR_DEBUG_NARROW_8=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
  echo "-- TSASR_DEBUG_NO_NARROW=1 (W streamed per cell tile, 8 producer warps: the kernel as before)"; TSASR_DEBUG_NO_NARROW=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
done
} > gpurun_out/s30_ab_narrow.txt 2>&1
cat gpurun_out/s30_ab_narrow.txt
