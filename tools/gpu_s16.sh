#!/bin/bash
# session 16: in-situ fit_batch with the N1 / N3 drop-ins (1 GPU): test + three timings (joint only, + projections, all)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_insitu_gpu.py -x -q > gpurun_out/s16_insitu_test.log 2>&1; echo "insitu test rc=$?"
tail -5 gpurun_out/s16_insitu_test.log
for w in joint all; do
  timeout 600 python tools/insitu_step.py --arm dropin --dropins $w --steps 8 --warmup 4 > gpurun_out/s16_insitu_$w.json 2> gpurun_out/s16_insitu_$w.err; echo "insitu $w rc=$?"
  python -c "
import json,sys
d=json.load(open('gpurun_out/s16_insitu_$w.json'))
print('$w', d['dropin']['ms_per_step'], d['dropin']['peak_mem_gib'], d['dropin']['losses'])"
done
timeout 600 python tools/insitu_step.py --arm both --dropins all --steps 8 --warmup 4 > gpurun_out/s16_insitu_both.json 2> gpurun_out/s16_insitu_both.err; echo "insitu both rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/s16_insitu_both.json'))
print(d['stock']['ms_per_step'], d['dropin']['ms_per_step'], d['speedup_fit_batch']); print(d['parity'])"
