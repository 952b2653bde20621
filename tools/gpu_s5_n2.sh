#!/bin/bash
# round-2 GPU session 5 (2 GPUs): the reference's DDP route end to end -- NCCL tests, bench weak + strong at N=2, in-situ at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py tests/test_hardening_gpu.py -m gpu -q -rs > gpurun_out/s5_pytest_n2.log 2>&1; echo "rc=$?" >> gpurun_out/s5_pytest_n2.log
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s5_bench_n2.json 2> gpurun_out/s5_bench_n2.err; echo "rc=$?" >> gpurun_out/s5_bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 6 --warmup 3 --global-batch 128 --sustain-s 0 > gpurun_out/s5_bench_n2_strong.json 2> gpurun_out/s5_bench_n2_strong.err; echo "rc=$?" >> gpurun_out/s5_bench_n2_strong.err
timeout 600 python bench.py --steps 6 --warmup 3 --global-batch 128 --sustain-s 0 --no-cpu-baseline --no-reference-gpu > gpurun_out/s5_bench_n1_strong.json 2> gpurun_out/s5_bench_n1_strong.err; echo "rc=$?" >> gpurun_out/s5_bench_n1_strong.err
timeout 900 $TR tools/insitu_step.py --steps 4 --warmup 2 > gpurun_out/s5_insitu_n2.json 2> gpurun_out/s5_insitu_n2.err; echo "rc=$?" >> gpurun_out/s5_insitu_n2.err
tail -4 gpurun_out/s5_pytest_n2.log
python - <<'PY'
import json
for f in ("s5_bench_n2", "s5_bench_n2_strong", "s5_bench_n1_strong"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "n_gpus", d["n_gpus"], d["scaling"], "B/gpu", d["config"]["B_per_gpu"], "value %.1f M  ms %.3f  e2e %.1f M" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6), d.get("ddp_check"))
    except Exception as ex:
        print(f, "failed", ex)
try:
    d = json.load(open("gpurun_out/s5_insitu_n2.json"))
    print("insitu n2: stock %.1f ms dropin %.1f ms speedup %.2f" % (d["stock"]["ms_per_step"], d["dropin"]["ms_per_step"], d["speedup_fit_batch"]), d["parity"])
except Exception as ex:
    print("insitu failed", ex)
PY
tail -3 gpurun_out/s5_insitu_n2.err
