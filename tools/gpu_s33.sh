#!/bin/bash
# session 33: narrow-vocabulary forward -- cost of MMA issue and of the A-block release commits; release granularity variants
mkdir -p gpurun_out
{
shape="16 400 240 640 29"
for skip in 0 512 1024; do
  echo "== TSASR_DEBUG_SKIP=$skip (512: A blocks released two at a time, 1024: all at the end of the cell tile)"
  TSASR_DEBUG_SKIP=$skip timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
  TSASR_DEBUG_SKIP=$skip TSASR_DEBUG_PROF=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | grep -E "mode=0|MMA issue" | tail -2
done
} > gpurun_out/s33_narrow_commit.txt 2>&1
cat gpurun_out/s33_narrow_commit.txt
