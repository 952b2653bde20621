#!/bin/bash
# round-2 GPU session 6: evidence of the final code -- bench, ncu launch list of the same command, ncu --set full of the hot kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_hardening_gpu.py -m gpu -q > gpurun_out/s6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s6_pytest.log
timeout 900 python tools/insitu_step.py --steps 6 --warmup 3 --deferred-check > gpurun_out/s6_insitu.json 2> gpurun_out/s6_insitu.err; echo "insitu rc=$?" >> gpurun_out/s6_insitu.err
timeout 600 python bench.py --config insitu --steps 3 --warmup 2 > gpurun_out/s6_bench_insitu.json 2> gpurun_out/s6_bench_insitu.err
timeout 600 python bench.py > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; echo "bench rc=$?" >> gpurun_out/s6_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain-s 0"
timeout 300 $CMD > gpurun_out/s6_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s6_launches.csv $CMD > gpurun_out/s6_ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/s6_plain2.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"joint_gemm_kernel|alpha_beta_kernel|dj_gemm_kernel|dw_gemm_kernel" -s 15 -c 5 -o gpurun_out/s6_prof -f $CMD > gpurun_out/s6_ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/s6_ncu_full.log
tail -3 gpurun_out/s6_pytest.log; tail -c 900 gpurun_out/s6_insitu.json; tail -3 gpurun_out/s6_bench.err; tail -3 gpurun_out/s6_ncu_list.log; tail -5 gpurun_out/s6_ncu_full.log; ls -la gpurun_out/s6_*
