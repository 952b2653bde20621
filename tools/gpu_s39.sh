#!/bin/bash
# session 39: does the CTA-pair protocol cost the narrow-vocabulary kernel?  single-CTA form (TSASR_DEBUG_NO_PAIR=1: W streamed, 8 producers)
# against the pair forms at the recipe's shape
mkdir -p gpurun_out
{
shape="16 400 240 640 29"
echo "== single CTA (cta_group::1), W streamed, 8 producers"; TSASR_DEBUG_NO_PAIR=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2
echo "== single CTA, role ablations: 1 epilogue only releases, 2 producers only arrive, 3 both"
for skip in 1 2 3; do TSASR_DEBUG_NO_PAIR=1 TSASR_DEBUG_SKIP=$skip timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2 | head -1; done
echo "== pairs, W streamed, 8 producers (TSASR_DEBUG_NO_NARROW=1)"; TSASR_DEBUG_NO_NARROW=1 timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2 | head -1
echo "== pairs, W resident, 16 producers (shipped)"; timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -2 | head -1
} > gpurun_out/s39_single_cta.txt 2>&1
cat gpurun_out/s39_single_cta.txt
