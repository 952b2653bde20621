#!/bin/bash
# session 14: N1 tests after the 2-CTA/SM change + ncu --set full capture of the three projection GEMMs (B128 shape)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py -q > gpurun_out/s14_linear_tests.log 2>&1; echo "linear tests rc=$?"
tail -5 gpurun_out/s14_linear_tests.log
timeout 120 python tools/ncu_linear_target.py && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_gemm -s 6 -c 3 -o gpurun_out/s14_linear -f \
    python tools/ncu_linear_target.py > gpurun_out/s14_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/s14_ncu.log
