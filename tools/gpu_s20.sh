#!/bin/bash
# session 20: full regression of the round-2 code with the N1 / N2 / N3 additions: whole GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/s20_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/s20_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s20_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/s20_smoke.log
timeout 600 python bench.py > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/s20_bench.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'steps', 'warmup')}, d['e2e']['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d.get('next_rows'), indent=0)[:1500])
print(d.get('clocks'))
PY
