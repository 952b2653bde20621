#!/bin/bash
# session 19 (2 GPUs): the in-situ fit_batch with EVERY drop-in (joint + loss, projections, prediction network) under the
# reference's per-module DDP, and the NCCL DDP test
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/insitu_step.py --steps 4 --warmup 2 --dropins all > gpurun_out/s19_insitu_n2.json 2> gpurun_out/s19_insitu_n2.err; echo "insitu rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/s19_insitu_n2.json"))
    print("insitu n2 (dropins=%s): stock %.1f ms dropin %.1f ms speedup %.2f" % (d["dropins"], d["stock"]["ms_per_step"], d["dropin"]["ms_per_step"], d["speedup_fit_batch"]), d["parity"])
except Exception as ex:
    print("insitu failed", ex)
PY
tail -3 gpurun_out/s19_insitu_n2.err
timeout 300 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -rs > gpurun_out/s19_pytest_n2.log 2>&1; echo "ddp test rc=$?"
tail -3 gpurun_out/s19_pytest_n2.log
