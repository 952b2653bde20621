"""Per-kernel timing of the hot path with CUDA events (development tool).

    python tools/bench_kernels.py [B T U H V] [--iters N]
Prints ms and TFLOP/s (2*M*H*V per GEMM-equivalent) for: joint fwd, DP, full backward."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsasr_b200 import _lib, ops  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dims", nargs="*", type=int, default=[16, 400, 100, 640, 1000])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ragged", action="store_true")
    a = ap.parse_args()
    B, T, U, H, V = a.dims
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().to(dev)
    dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().to(dev)
    W = ((torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5).bfloat16().to(dev)
    b = ((torch.rand(V, generator=g) * 2 - 1) / H ** 0.5).to(dev)
    tg = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).to(dev)
    ll = torch.full((B,), T, dtype=torch.int32)
    tl = torch.full((B,), U - 1, dtype=torch.int32)
    if a.ragged:
        ll = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, dtype=torch.int32)
        tl = torch.randint(int(0.4 * U), U, (B,), generator=g, dtype=torch.int32)
        ll[0], tl[0] = T, U - 1
    live = int(((ll.long() + 15) // 16 * 16 * ((tl.long() + 1 + 7) // 8 * 8)).sum()) if (T, U) == (400, 100) else None
    print(f"lengths: T_b={ll.tolist()} labels={tl.tolist()}; cells in live 16x8 tiles: {live} of {B * T * U}")
    ll, tl = ll.to(dev), tl.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gemm = 2.0 * B * T * U * H * V
    dcost = torch.ones(B, device=dev)

    med, best = timeit(lambda: ops.joint_fwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01), a.iters, flush)
    print(f"joint_fwd   : {med:8.3f} ms (best {best:.3f})  {gemm / med / 1e9:8.1f} TFLOP/s")
    lat2, logz = ops.joint_fwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01)
    med, best = timeit(lambda: ops.alpha_beta(lat2, ll, tl, B, T, U), a.iters, flush)
    print(f"alpha_beta  : {med:8.3f} ms (best {best:.3f})")
    alpha, beta, cost, _, _ = ops.alpha_beta(lat2, ll, tl, B, T, U)
    med, best = timeit(lambda: ops.joint_bwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01, lat2, logz, alpha, beta, cost, dcost),
                       a.iters, flush)
    print(f"joint_bwd   : {med:8.3f} ms (best {best:.3f})  {3 * gemm / med / 1e9:8.1f} TFLOP/s (3 GEMM-equivalents executed)")
    print(f"cells/s (fwd+dp+bwd medians summed) launches={_lib.launch_count()}")


if __name__ == "__main__":
    main()
