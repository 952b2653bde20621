#!/bin/bash
# A/B of epilogue experiments (three-input max tree, bias straight from global memory) on one box, alternating, config 2 and H=256
mkdir -p gpurun_out
D=$PWD/tsasr_b200
for v in exp_MAX3 exp_LDGBIAS exp_MAX3_LDGBIAS; do
  TSASR_B200_LIB=$D/libtsasr_b200_$v.so timeout 600 python -m pytest tests/test_joint_gpu.py -q -m gpu -x > gpurun_out/s11_pytest_$v.log 2>&1; echo "$v pytest rc=$?" >> gpurun_out/s11_ab.log
done
for rep in 1 2 3; do
for shape in "16 400 100 640 1000" "8 750 200 640 5000"; do
  echo "== shape $shape (rep $rep)" >> gpurun_out/s11_ab.log
  for v in exp exp_MAX3 exp_LDGBIAS exp_MAX3_LDGBIAS; do
    TSASR_B200_LIB=$D/libtsasr_b200_$v.so timeout 200 python tools/time_fwd.py $shape 2>&1 | sed "s|$D/libtsasr_b200_||" >> gpurun_out/s11_ab.log
  done
done
done
cat gpurun_out/s11_ab.log
