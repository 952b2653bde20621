#!/bin/bash
# session 34: narrow vocabularies -- A blocks handed over in two groups per cell tile; parity suites; where the gradient pass spends its time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_joint_gpu.py tests/test_hardening_gpu.py -x -q > gpurun_out/s34_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s34_tests.log
{
for shape in "16 400 240 640 29" "16 400 100 256 64"; do
  echo "== shape $shape"
  echo "-- default (resident W, 16 producer warps, two hand-overs per cell tile)"; timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
  echo "-- TSASR_DEBUG_NARROW_8=1 (resident W, 8 producer warps, one hand-over per k-block)"; TSASR_DEBUG_NARROW_8=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
done
shape="16 400 240 640 29"
echo "== tile pruning off (TSASR_PRUNE_LOG2_EPS=0)"; TSASR_PRUNE_LOG2_EPS=0 timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -1
echo "== no programmatic dependent launch (TSASR_DEBUG_NO_PDL=1)"; TSASR_DEBUG_NO_PDL=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
echo "== MMA-lane profile"; TSASR_DEBUG_PROF=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | grep -E "mode=" | awk 'NR%9==1' | tail -4
} > gpurun_out/s34_narrow_groups.txt 2>&1
cat gpurun_out/s34_narrow_groups.txt
