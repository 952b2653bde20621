#!/bin/bash
# session 17: batched beam search (tests + timing against the real reference searcher)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py -x -q > gpurun_out/s17_decode_tests.log 2>&1; echo "decode tests rc=$?"
tail -4 gpurun_out/s17_decode_tests.log
timeout 900 python tools/bench_beam.py > gpurun_out/s17_bench_beam.txt 2> gpurun_out/s17_bench_beam.err; echo "bench beam rc=$?"
cat gpurun_out/s17_bench_beam.txt; tail -3 gpurun_out/s17_bench_beam.err
