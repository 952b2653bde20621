#!/bin/bash
# session 18: beam-search timing against the real reference searcher (rows time-limited), bench.py with the next_rows leg
mkdir -p gpurun_out
timeout 420 python tools/bench_beam.py > gpurun_out/s18_bench_beam.txt 2> gpurun_out/s18_bench_beam.err; echo "bench beam rc=$?"
cat gpurun_out/s18_bench_beam.txt; tail -3 gpurun_out/s18_bench_beam.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s18_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'])
print(json.dumps(d.get('next_rows'), indent=1))
print(d['roofline'])
PY
tail -3 gpurun_out/s18_bench.err
