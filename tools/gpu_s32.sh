#!/bin/bash
# session 32: what paces the narrow-vocabulary forward -- MMA-lane wait breakdown and role ablations (TSASR_DEBUG_SKIP bits)
mkdir -p gpurun_out
{
shape="16 400 240 640 29"
echo "== MMA-lane profile (TSASR_DEBUG_PROF=1), forward launches"
TSASR_DEBUG_PROF=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | grep -E "mode=0" | tail -3
echo "== MMA-lane profile, gradient pass"
TSASR_DEBUG_PROF=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | grep -E "mode=1|bwd kernels" | tail -3
for skip in 0 1 2 32 3 34 35; do
  echo "== TSASR_DEBUG_SKIP=$skip (1: epilogue only releases, 2: producers only arrive, 32: no MMAs issued)"
  TSASR_DEBUG_SKIP=$skip timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
done
} > gpurun_out/s32_narrow_ablation.txt 2>&1
cat gpurun_out/s32_narrow_ablation.txt
