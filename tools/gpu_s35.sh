#!/bin/bash
# session 35: ncu --set full of the narrow-vocabulary forward / gradient pass and of the backward GEMMs at the recipe's shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_joint_gpu.py tests/test_hardening_gpu.py -x -q > gpurun_out/s35_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s35_tests.log
timeout 120 python tools/one_step.py 16 400 240 640 29 2 && \
timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 7 --launch-count 12 -o gpurun_out/s35_recipe -f python tools/one_step.py 16 400 240 640 29 2 > gpurun_out/s35_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/s35_ncu.log; ls -la gpurun_out/s35_recipe.ncu-rep
