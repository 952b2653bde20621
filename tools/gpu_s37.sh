#!/bin/bash
# session 37: regression of the final code -- every GPU test, smoke, the bench line (headline + recipe shape), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s37_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s37_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/s37_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s37_smoke.log
timeout 600 python bench.py > gpurun_out/s37_bench.json 2> gpurun_out/s37_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/s37_bench.json
timeout 300 python bench.py --shape recipe --no-cpu-baseline --no-reference-gpu --sustain-s 0 > gpurun_out/s37_bench_recipe.json 2> gpurun_out/s37_bench_recipe.err; echo "bench recipe rc=$?"; cut -c1-400 gpurun_out/s37_bench_recipe.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain-s 0"
timeout 300 $CMD > gpurun_out/s37_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s37_launches.csv $CMD > gpurun_out/s37_ncu_list.log 2>&1
echo "launch list rc=$?"
