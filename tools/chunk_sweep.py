"""Backward chunk-size sweep (evidence tool): does keeping the operand images of a chunk L2-resident pay?

The backward stages bf16 dlogits + J images through the workspace; with the default 3 GiB chunk the whole batch is one
chunk at config 2 (2.16 GB of images: GRAD writes them to HBM, dj / dw read them back).  Smaller chunks keep the images
in the 126 MB L2 between the three kernels, at the price of more launches (4 per chunk) and of partially filled last
waves (74 CTA pairs x 2 tiles per wave).  python tools/chunk_sweep.py [prune_log2_eps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsasr_b200 import _lib, ops  # noqa: E402

B, T, U, H, V = 16, 400, 100, 640, 1000
dev = torch.device("cuda:0")
eps = float(sys.argv[1]) if len(sys.argv) > 1 else -30.0
g = torch.Generator().manual_seed(0)
enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().to(dev)
dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().to(dev)
W = ((torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5).bfloat16().to(dev)
b = ((torch.rand(V, generator=g) * 2 - 1) / H ** 0.5).to(dev)
tg = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).to(dev)
ll = torch.full((B,), T, dtype=torch.int32, device=dev)
tl = torch.full((B,), U - 1, dtype=torch.int32, device=dev)
dcost = torch.full((B,), 1.0 / B, device=dev)
lat2, logz = ops.joint_fwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01)
alpha, beta, cost, _, _ = ops.alpha_beta(lat2, ll, tl, B, T, U)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tile_bytes = (16 + 10) * 16384
print(f"config 2 (B={B} T={T} U={U} V={V} H={H}), prune_log2_eps={eps}; 5200 tiles of {tile_bytes // 1024} KB of images each")
print(f"{'chunk tiles':>12s} {'chunks':>7s} {'images MB/chunk':>16s} {'bwd ms':>8s}   per-kernel ms (GRAD, dj, dw, folds+prune)")
for tiles in (5200, 2600, 1300, 650, 450, 300, 150):
    cells = tiles * 128

    def bwd():
        return ops.joint_bwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01, lat2, logz, alpha, beta, cost, dcost,
                             max_chunk_cells=cells, prune_log2_eps=eps)
    for _ in range(2):
        bwd()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        bwd()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    _lib.kernel_timing(True)
    flush.zero_()
    bwd()
    torch.cuda.synchronize()
    km = _lib.kernel_timings()
    _lib.kernel_timing(False)
    n_chunks = km.get("dw_gemm_kernel", (0, 1))[1]
    grad = km.get("joint_gemm_kernel<GRAD>", (0, 0))[0]
    dj = km.get("dj_gemm_kernel", (0, 0))[0]
    dw = km.get("dw_gemm_kernel", (0, 0))[0]
    rest = sum(v[0] for k, v in km.items()) - grad - dj - dw
    print(f"{tiles:12d} {n_chunks:7d} {tiles * tile_bytes / 2**20:16.0f} {sorted(ts)[len(ts) // 2]:8.3f}   {grad:.3f}  {dj:.3f}  {dw:.3f}  {rest:.3f}")
