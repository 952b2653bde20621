#!/bin/bash
# session 22 (8 GPUs): BASELINE configs[4] -- the full train step in situ on 8 x B200 with every drop-in, reference DDP
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563"
timeout 500 $TR tools/insitu_step.py --steps 6 --warmup 3 --dropins all > gpurun_out/s22_insitu_n8.json 2> gpurun_out/s22_insitu_n8.err; echo "insitu rc=$?"
python - <<'PY'
import json
for line in open("gpurun_out/s22_insitu_n8.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("insitu n8 (dropins=%s): stock %.1f ms dropin %.1f ms speedup %.2f  mem %.1f -> %.1f GiB" % (d["dropins"], d["stock"]["ms_per_step"], d["dropin"]["ms_per_step"], d["speedup_fit_batch"], d["stock"]["peak_mem_gib"], d["dropin"]["peak_mem_gib"]))
        print({k: (round(v, 6) if isinstance(v, float) else v) for k, v in d["parity"].items() if not k.startswith("losses")})
PY
tail -2 gpurun_out/s22_insitu_n8.err
