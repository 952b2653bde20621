#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rf --durations=6 > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
timeout 300 python tools/bench_greedy.py > gpurun_out/s4_greedy.log 2>&1; echo "rc=$?" >> gpurun_out/s4_greedy.log
tail -15 gpurun_out/s4_pytest.log; cat gpurun_out/s4_greedy.log
