#!/bin/bash
# session 42: last check of the in-tree build (same sources as session 37): smoke, the joint / hardening / lattice suites, a short bench line
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/s42_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s42_smoke.log
timeout 600 python -m pytest tests/test_joint_gpu.py tests/test_hardening_gpu.py tests/test_lattice_gpu.py tests/test_linear_gpu.py tests/test_predictor_gpu.py tests/test_decode_gpu.py -q > gpurun_out/s42_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/s42_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain-s 0 > gpurun_out/s42_bench.json 2> gpurun_out/s42_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/s42_bench.json
