#!/bin/bash
# session 40: dJ epilogue -- accumulator released before the partial-row stores, next unit's scalars staged through shared memory with cp.async,
# 16-column TMEM loads; backward parity suites, timing at the recipe's shape, config 2 and a V = 5000 shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_joint_gpu.py tests/test_hardening_gpu.py -x -q > gpurun_out/s40_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s40_tests.log
{
for shape in "16 400 240 640 29" "16 400 100 640 1000" "4 400 100 640 5000" "16 400 100 256 1000"; do
  echo "== shape $shape"; timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -1
done
} > gpurun_out/s40_dj_epilogue.txt 2>&1
cat gpurun_out/s40_dj_epilogue.txt
