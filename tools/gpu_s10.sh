#!/bin/bash
# A/B of the producer change (one enc load + two k-blocks of prefetch) against the previous build, same box, alternating
mkdir -p gpurun_out
BASE=$PWD/tsasr_b200/libtsasr_b200_base.so
timeout 900 python -m pytest tests/test_joint_gpu.py tests/test_lattice_gpu.py tests/test_fullsize_gpu.py -q -m gpu > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s10_pytest.log
for rep in 1 2; do
for shape in "16 400 100 640 1000" "16 400 240 640 29" "8 750 200 640 5000" "16 400 100 256 1000"; do
  echo "== shape $shape (rep $rep)" >> gpurun_out/s10_ab.log
  TSASR_B200_LIB=$BASE timeout 200 python tools/time_fwd.py $shape >> gpurun_out/s10_ab.log 2>&1
  timeout 200 python tools/time_fwd.py $shape >> gpurun_out/s10_ab.log 2>&1
done
done
tail -3 gpurun_out/s10_pytest.log; cat gpurun_out/s10_ab.log
