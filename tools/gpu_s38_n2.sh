#!/bin/bash
# session 38 (2 GPUs): the two-GPU tests and the N=2 bench line (reference DDP route) on the final code
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 400 python -m pytest tests/test_ddp_nccl_gpu.py tests/test_hardening_gpu.py -m gpu -q -rs -k "ddp or two or device or nccl" > gpurun_out/s38_pytest_n2.log 2>&1; echo "two-GPU tests rc=$?"; tail -3 gpurun_out/s38_pytest_n2.log
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain-s 0 > gpurun_out/s38_bench_n2.json 2> gpurun_out/s38_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/s38_bench_n2.json").read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], d["scaling"], "value %.1f M  ms %.3f  e2e %.1f M" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6), d.get("ddp_check"))
except Exception as ex:
    print("bench n2 failed", ex)
PY
