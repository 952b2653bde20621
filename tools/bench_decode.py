"""Decode-time joint step: fused kernel vs the reference's eager chain (development / evidence tool).

    python tools/bench_decode.py
Per (B, H, V): microseconds per step, device time (CUDA events over 300 back-to-back steps) and host wall time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsasr_b200  # noqa: E402
from tsasr_b200 import decode  # noqa: E402
from oracle.greedy_decode import eager_joint_step  # noqa: E402

dev = torch.device("cuda:0")


def bench(fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n, (time.perf_counter() - t0) * 1e6 / n


for B, H, V in ((1, 640, 1000), (16, 640, 1000), (32, 640, 1000), (16, 640, 5000)):
    head = torch.nn.Linear(H, V).to(dev)
    tjoint = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
    sm = torch.nn.LogSoftmax(dim=-1)
    h_i, out_pn = torch.randn(B, 1, 1, H, device=dev), torch.randn(B, 1, 1, H, device=dev)
    eager = eager_joint_step(tjoint, [head], sm)
    fused = decode.fused_joint_forward_step(tjoint, [head], sm)
    err = (eager(h_i, out_pn) - fused(h_i, out_pn)).abs().max().item()
    de, we = bench(lambda: eager(h_i, out_pn))
    df, wf = bench(lambda: fused(h_i, out_pn))
    tsasr_b200._lib.kernel_timings()
    tsasr_b200._lib.kernel_timing(True)
    for _ in range(50):
        fused(h_i, out_pn)
    torch.cuda.synchronize()
    k = tsasr_b200._lib.kernel_timings()["joint_decode_step"]
    tsasr_b200._lib.kernel_timing(False)
    print(f"B={B:3d} H={H} V={V}: eager chain {de:6.1f} us/step (host {we:6.1f}) | fused {df:6.1f} us/step (host {wf:6.1f}), "
          f"kernels alone {k[0] / k[1] * 1e3:5.1f} us | max |dlogp| = {err:.1e}")
