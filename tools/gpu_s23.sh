#!/bin/bash
# session 23: evidence of the FINAL round-2 code -- ncu launch list of the bench command, ncu --set full of the five hot kernels,
# launch list of one in-situ-style step of the N1 / N3 kernels (single pass, no replay)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain-s 0"
timeout 300 $CMD > gpurun_out/s23_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s23_launches.csv $CMD > gpurun_out/s23_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 300 $CMD > gpurun_out/s23_plain2.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"joint_gemm_kernel|alpha_beta_kernel|dj_gemm_kernel|dw_gemm_kernel" -s 15 -c 5 -o gpurun_out/s23_prof -f $CMD > gpurun_out/s23_ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -2 gpurun_out/s23_ncu_full.log
CMD2="python tools/bench_predictor.py --iters 2"
timeout 200 $CMD2 > gpurun_out/s23_pred_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lstm_seq|onehot_dw|linear_gemm|linear_fold" -c 60 --csv --log-file gpurun_out/s23_launches_predictor.csv $CMD2 > gpurun_out/s23_ncu_pred.log 2>&1
echo "predictor launch list rc=$?"
ls -la gpurun_out/s23_*
