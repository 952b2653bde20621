#!/bin/bash
# round-2 GPU session 3: full GPU suite on the single-call forward + fused clamp + on-device greedy; host timeline; bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rf --durations=6 > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
timeout 300 python tools/profile_e2e_host.py > gpurun_out/s3_host_profile.log 2>&1; echo "rc=$?" >> gpurun_out/s3_host_profile.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?" >> gpurun_out/s3_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --shape recipe > gpurun_out/s3_bench_recipe.json 2> gpurun_out/s3_bench_recipe.err
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --ragged > gpurun_out/s3_bench_ragged.json 2> gpurun_out/s3_bench_ragged.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1
tail -12 gpurun_out/s3_pytest.log; head -20 gpurun_out/s3_host_profile.log; tail -2 gpurun_out/s3_bench.err; tail -1 gpurun_out/s3_smoke.log
python - <<'PY'
import json
for f in ("s3_bench", "s3_bench_recipe", "s3_bench_ragged"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.1f M  ms %.3f  e2e %.1f M (%.3f ms)  sustained %.1f M  fwd frac_burst %.3f  dense %.3f ms" % (
            d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["sustained"]["value"] / 1e6,
            d["roofline"]["frac_burst"], d.get("dense_backward", {}).get("ms_per_step", 0)))
    except Exception as ex:
        print(f, "failed", ex)
PY
