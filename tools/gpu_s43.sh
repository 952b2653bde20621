#!/bin/bash
# session 43: the recipe AS SHIPPED in situ -- 29 characters, 239 labels (U = 240) on 16 s of audio (T = 400), B = 16, one GPU:
# full fit_batch of the real TSASR Brain, stock modules vs every drop-in
mkdir -p gpurun_out
timeout 600 python tools/insitu_step.py --vocab 29 --labels 239 --steps 6 --warmup 3 --dropins all > gpurun_out/s43_insitu_v29.json 2> gpurun_out/s43_insitu_v29.err; echo "insitu rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/s43_insitu_v29.json").read().strip().splitlines()[-1])
    print("insitu V=29 (dropins=%s): stock %.1f ms dropin %.1f ms speedup %.2f" % (d["dropins"], d["stock"]["ms_per_step"], d["dropin"]["ms_per_step"], d["speedup_fit_batch"]))
    print({k: d[k] for k in d if k in ("parity", "peak_mem_gib", "config")})
    print("stock", {k: v for k, v in d["stock"].items() if k != "losses"}, "dropin", {k: v for k, v in d["dropin"].items() if k != "losses"})
except Exception as ex:
    print("insitu failed", ex)
PY
tail -3 gpurun_out/s43_insitu_v29.err
