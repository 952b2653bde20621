#!/bin/bash
# round-2 GPU session 2: new parity / hardening / in-situ tests, the new bench legs, the in-situ step, an ncu launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_parity_tight_gpu.py tests/test_hardening_gpu.py tests/test_insitu_gpu.py -m gpu -q -rA --durations=8 > gpurun_out/s2_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_parity_tight_gpu.py --deselect tests/test_hardening_gpu.py --deselect tests/test_insitu_gpu.py > gpurun_out/s2_pytest_old.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest_old.log
timeout 600 python tools/explore_grad_error.py > gpurun_out/s2_explore.log 2>&1; echo "rc=$?" >> gpurun_out/s2_explore.log
timeout 600 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?" >> gpurun_out/s2_bench.err
timeout 900 python tools/insitu_step.py --steps 6 --warmup 3 > gpurun_out/s2_insitu.json 2> gpurun_out/s2_insitu.err; echo "insitu rc=$?" >> gpurun_out/s2_insitu.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain-s 0"
timeout 300 $CMD > gpurun_out/s2_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2_launches.csv $CMD > gpurun_out/s2_ncu.log 2>&1
grep -E "passed|failed|rc=" gpurun_out/s2_pytest_new.log | tail -5; grep -E "^(FAILED|ERROR)" gpurun_out/s2_pytest_new.log | head -20
tail -2 gpurun_out/s2_pytest_old.log; tail -3 gpurun_out/s2_bench.err; tail -3 gpurun_out/s2_insitu.err; head -c 1500 gpurun_out/s2_insitu.json
