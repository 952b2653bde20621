#!/bin/bash
# session 24: A/B of the lazy softmax reference point in the forward epilogue (-DTSASR_LAZY_RESCALE build) against the shipped kernel
mkdir -p gpurun_out
L=tsasr_b200/libtsasr_b200_lazy.so
for i in 1 2 3; do
  timeout 120 python tools/time_fwd.py 2>/dev/null | head -1
  TSASR_B200_LIB=$PWD/$L timeout 120 python tools/time_fwd.py 2>/dev/null | head -1
done | tee gpurun_out/s24_ab_lazy.txt
TSASR_B200_LIB=$PWD/$L timeout 600 python -m pytest tests/test_joint_gpu.py -x -q > gpurun_out/s24_lazy_tests.log 2>&1; echo "lazy parity rc=$?"; tail -2 gpurun_out/s24_lazy_tests.log
for shape in "16 400 240 640 29" "8 750 200 640 5000"; do
  timeout 120 python tools/time_fwd.py $shape 2>/dev/null | head -1
  TSASR_B200_LIB=$PWD/$L timeout 120 python tools/time_fwd.py $shape 2>/dev/null | head -1
done | tee -a gpurun_out/s24_ab_lazy.txt
