"""Backward tile pruning: active tiles, backward time and deviation from the unpruned gradients as a function of the
threshold (evidence tool).  python tools/prune_sweep.py  -> config 2, full-length and ragged batches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsasr_b200 import ops  # noqa: E402

B, T, U, H, V = 16, 400, 100, 640, 1000
dev = torch.device("cuda:0")


def run(ragged):
    g = torch.Generator().manual_seed(0)
    enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().to(dev)
    dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().to(dev)
    W = ((torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5).bfloat16().to(dev)
    b = ((torch.rand(V, generator=g) * 2 - 1) / H ** 0.5).to(dev)
    tg = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).to(dev)
    ll = torch.full((B,), T, dtype=torch.int32)
    tl = torch.full((B,), U - 1, dtype=torch.int32)
    if ragged:
        ll = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, dtype=torch.int32)
        tl = torch.randint(int(0.4 * U), U, (B,), generator=g, dtype=torch.int32)
        ll[0], tl[0] = T, U - 1
    ll, tl = ll.to(dev), tl.to(dev)
    dcost = torch.full((B,), 1.0 / B, device=dev)
    lat2, logz = ops.joint_fwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01)
    alpha, beta, cost, _, _ = ops.alpha_beta(lat2, ll, tl, B, T, U)

    def bwd(eps):
        return ops.joint_bwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01, lat2, logz, alpha, beta, cost, dcost, prune_log2_eps=eps)

    ref = [x.clone() for x in bwd(0.0)]
    print(f"{'ragged' if ragged else 'full-length'} batch, B={B} T={T} U={U} V={V} H={H}")
    print(f"{'log2 eps':>9s} {'active/live tiles':>18s} {'bwd ms':>8s}   max |g - g_dense| / max |g_dense|  (d_enc, d_dec, dW, db)")
    for eps in (0.0, -60.0, -40.0, -30.0, -24.0, -20.0, -16.0, -10.0):
        for _ in range(2):
            bwd(eps)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            out = bwd(eps)
        e.record()
        torch.cuda.synchronize()
        act, live = ops.last_backward_tile_stats(dev) if eps < 0 else (None, None)
        errs = [((o - r).abs().max() / r.abs().max()).item() for o, r in zip(out, ref)]
        frac = f"{act}/{live}" if act is not None else "all"
        print(f"{eps:9.0f} {frac:>18s} {s.elapsed_time(e) / 5:8.3f}   " + "  ".join(f"{x:.1e}" for x in errs))


run(False)
run(True)
