#!/bin/bash
# round-2 GPU session 8: final regression exactly as the driver runs it (pytest -m gpu -x, smoke, both bench arms)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s8_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s8_bench_ref.json 2> gpurun_out/s8_bench_ref.err; echo "rc=$?" >> gpurun_out/s8_bench_ref.err
timeout 600 python bench.py > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; echo "rc=$?" >> gpurun_out/s8_bench.err
timeout 300 python bench.py --shape config4 --no-cpu-baseline --no-reference-gpu --steps 5 > gpurun_out/s8_bench_config4.json 2> gpurun_out/s8_bench_config4.err; echo "rc=$?" >> gpurun_out/s8_bench_config4.err
tail -4 gpurun_out/s8_pytest.log; tail -2 gpurun_out/s8_smoke.log; tail -1 gpurun_out/s8_bench_ref.err; tail -1 gpurun_out/s8_bench.err; tail -1 gpurun_out/s8_bench_config4.err
python - <<'PY'
import json
for f in ("s8_bench", "s8_bench_config4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.1f M  ms %.3f  e2e %.3f ms  sustained %.3f ms  dense %.3f ms  fwd frac_burst %.3f  step frac_burst %.3f" % (
            d["value"] / 1e6, d["ms_per_step"], d["e2e"]["ms_per_step"], d["sustained"]["ms_per_step"], d["dense_backward"]["ms_per_step"],
            d["roofline"]["frac_burst"], d["roofline"]["step"]["frac_burst"]))
    except Exception as ex:
        print(f, "failed", ex)
PY
