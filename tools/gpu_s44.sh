#!/bin/bash
# session 44: the in-situ test with both vocabularies (256 and the recipe's 29 characters)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_insitu_gpu.py -q > gpurun_out/s44_insitu_test.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/s44_insitu_test.log
