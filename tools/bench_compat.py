"""HBM-bound kernels of the path against the measured copy bandwidth (development / evidence tool).

    python tools/bench_compat.py            # compat path at config-2 shape + DP at growing batch
Per kernel: CUDA-event time from the library's own launch brackets (tsasr_kernel_timing_enable), algorithmic
bytes and the fraction of MEASURED_PEAKS.json hbm_gbs.  L2 is flushed between launches."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tsasr_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6650.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    _lib.kernel_timings()
    _lib.kernel_timing(True)
    for _ in range(iters):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    t = _lib.kernel_timings()
    _lib.kernel_timing(False)
    return {k: v[0] / v[1] for k, v in t.items()}


def line(name, ms, nbytes):
    gbs = nbytes / (ms / 1e3) / 1e9
    print(f"{name:44s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs / HBM:6.3f} of measured HBM ({HBM:.0f} GB/s)")


def compat(B, T, U, V, dtype=torch.float32):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(B, T, U, V, generator=g).to(dev).to(dtype)
    tg = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).to(dev)
    ll = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), U - 1, dtype=torch.int32, device=dev)
    M, es = B * T * U, logits.element_size()
    t = timed(lambda: ops.logits_to_lattice(logits, tg, ll, tl, 0))
    line(f"logits_to_lattice_kernel {tuple(logits.shape)} {str(dtype)[6:]}", t["logits_to_lattice_kernel"], es * M * V + 12 * M)
    lat2, den = ops.logits_to_lattice(logits, tg, ll, tl, 0)
    alpha, beta, cost, _, _ = ops.alpha_beta(lat2, ll, tl, B, T, U)
    dcost = torch.ones(B, device=dev)
    t = timed(lambda: ops.logits_grad(logits, tg, ll, tl, 0, lat2, den, alpha, beta, cost, dcost))
    line(f"logits_grad_kernel {tuple(logits.shape)} {str(dtype)[6:]}", t["logits_grad_kernel"], 2 * es * M * V + 20 * M)


def dp(B, T, U):
    n = ops.lattice_elems(B, T, U)
    lat2 = -torch.rand(n, 2, device=dev) * 5 - 0.1
    ll = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), U - 1, dtype=torch.int32, device=dev)
    t = timed(lambda: ops.alpha_beta(lat2, ll, tl, B, T, U))
    ms = t["alpha_beta_kernel"]
    line(f"alpha_beta_kernel B={B} T={T} U={U} ({ms * 1e6 / (T + U - 1):.0f} ns/diagonal)", ms, 24.0 * B * T * U)


if __name__ == "__main__":
    compat(16, 400, 100, 1000)
    compat(16, 400, 100, 1000, torch.float16)
    compat(4, 200, 40, 1000)
    for B in (16, 128, 512, 2048):
        dp(B, 400, 100)
