#!/bin/bash
# session 41: dJ epilogue A/B -- three builds in one call: base (committed), staged scalars + late release, staged scalars + early release
mkdir -p gpurun_out
{
for shape in "16 400 240 640 29" "16 400 100 640 1000"; do
  for lib in base late ""; do
    so=tsasr_b200/libtsasr_b200${lib:+_$lib}.so
    echo "== shape $shape  build ${lib:-early}"; TSASR_B200_LIB=$PWD/$so timeout 120 python tools/time_fwd.py $shape 2>&1 | tail -1
  done
done
} > gpurun_out/s41_dj_ab.txt 2>&1
cat gpurun_out/s41_dj_ab.txt
