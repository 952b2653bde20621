#!/bin/bash
# session 45: ncu --set full of the five hot kernels of the bench step on the FINAL code (kernel names with the third template parameter)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain-s 0"
timeout 300 $CMD > gpurun_out/s45_plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"joint_gemm_kernel|alpha_beta_kernel|dj_gemm_kernel|dw_gemm_kernel" -s 15 -c 5 -o gpurun_out/s45_prof -f $CMD > gpurun_out/s45_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/s45_ncu_full.log; ls -la gpurun_out/s45_prof.ncu-rep
