#!/bin/bash
# session 31: what paces the narrow-vocabulary form -- MMA-lane wait breakdown (TSASR_DEBUG_PROF) of forward and gradient pass
mkdir -p gpurun_out
{
for shape in "16 400 240 640 29"; do
  echo "== shape $shape, default"; TSASR_DEBUG_PROF=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | grep -E "tsasr prof|joint_fwd|bwd kernels" | sort | uniq -c | sort -rn | head -12
  echo "== shape $shape, TSASR_DEBUG_NARROW_8=1"; TSASR_DEBUG_NARROW_8=1 TSASR_DEBUG_PROF=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | grep -E "tsasr prof|joint_fwd|bwd kernels" | sort | uniq -c | sort -rn | head -12
done
} > gpurun_out/s31_prof_narrow.txt 2>&1
cat gpurun_out/s31_prof_narrow.txt
