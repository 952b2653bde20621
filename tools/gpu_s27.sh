#!/bin/bash
# session 27: reduce_dpre_kernel with staged tile flags and eight partial rows in flight -- parity suites + kernel timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_joint_gpu.py tests/test_parity_tight_gpu.py tests/test_fullsize_gpu.py tests/test_hardening_gpu.py -x -q > gpurun_out/s27_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/s27_tests.log
timeout 300 python tools/time_fwd.py 2>/dev/null | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/s27_bench.json 2> gpurun_out/s27_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/s27_bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
for k in d["kernels"]:
    if "reduce" in k["kernel"] or "dj" in k["kernel"]:
        print(k)
PY
