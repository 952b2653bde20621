#!/bin/bash
# round-2 GPU session 7 (8 GPUs): weak and strong scaling points through the reference's DDP route, in-situ at N=8
mkdir -p gpurun_out
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/s7_bench_n${N}.json 2> gpurun_out/s7_bench_n${N}.err; echo "rc=$?" >> gpurun_out/s7_bench_n${N}.err
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --global-batch 128 --sustain-s 0 > gpurun_out/s7_bench_n${N}_strong.json 2> gpurun_out/s7_bench_n${N}_strong.err; echo "rc=$?" >> gpurun_out/s7_bench_n${N}_strong.err
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 $TR tools/insitu_step.py --steps 4 --warmup 2 > gpurun_out/s7_insitu_n8.json 2> gpurun_out/s7_insitu_n8.err; echo "rc=$?" >> gpurun_out/s7_insitu_n8.err
for N in 2 1; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --global-batch 128 --sustain-s 0 --no-cpu-baseline --no-reference-gpu > gpurun_out/s7_bench_n${N}_strong.json 2> gpurun_out/s7_bench_n${N}_strong.err; echo "rc=$?" >> gpurun_out/s7_bench_n${N}_strong.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/s7_bench_n1.json 2> gpurun_out/s7_bench_n1.err
python - <<'PY'
import json
def load(f):
    try:
        for line in open(f):
            if line.strip().startswith("{"): return json.loads(line)
    except Exception as ex:
        return None
for f in ("s7_bench_n1","s7_bench_n4","s7_bench_n8","s7_bench_n1_strong","s7_bench_n2_strong","s7_bench_n4_strong","s7_bench_n8_strong"):
    d = load(f"gpurun_out/{f}.json")
    if d: print(f, d["n_gpus"], d["scaling"], "B/gpu", d["config"]["B_per_gpu"], "value %.1f M ms %.3f e2e %.1f M (%.3f ms)" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]), (d.get("ddp_check") or {}).get("head_grad_vs_rank_average_of_local_grads_max_rel"))
    else: print(f, "failed")
d = load("gpurun_out/s7_insitu_n8.json")
if d: print("insitu n8 stock %.1f dropin %.1f speedup %.2f loss err %.2e" % (d["stock"]["ms_per_step"], d["dropin"]["ms_per_step"], d["speedup_fit_batch"], d["parity"]["loss_rel_err_max"]))
PY
tail -2 gpurun_out/s7_*.err | tail -30
