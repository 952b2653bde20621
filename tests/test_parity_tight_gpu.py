"""Tight parity tests of the fused backward (round-2 additions; all calls go through the C ABI).

Three layers, so that a failure says WHERE the deviation is:

1. dlogits, element by element: the bf16 operand images the GRAD pass leaves in the workspace are decoded and compared
   with the reference's own dlogits (autograd of torchaudio's rnnt_loss through the head Linear): north_star's 1e-3
   max-abs plus the bf16 rounding of the value itself (2^-8 relative), with tile pruning ON, at a ragged config-2
   utterance pair; cells of pruned tiles must be negligible in the REFERENCE's dlogits.
2. The two backward GEMMs and the broadcast-sum reductions, exactly: a float64 redo on the decoded images must agree
   with dW, db, d_enc, d_dec to fp32-accumulation accuracy, elementwise, relative to the sum of |terms| of that
   element (1e-5) -- a dropped or double-counted (tile, v-block, split-K) item would show up here at O(1e-2 .. 1).
3. End to end against the reference chain, elementwise: |got - ref| <= K_SIGMA * sigma + (R_SYS + r_lat) * sum|terms|,
   where sigma is the standard deviation the bf16 rounding of dlogits (uniform, 2^-9 relative, independent per element)
   induces in THAT output element (computed from the reference's dlogits / joint tensors with a GEMM of squares), R_SYS
   bounds the common-mode error of the approximate exp2 / log2 chain, and r_lat is the floor the fp32 lattice imposes on
   ANY fp32 implementation, the reference included (measured against float64 by tools/fp32_noise_floor.py).  No tolerance
   is relative to the largest entry of the tensor, so small-magnitude regions (short utterances, padded tiles, the
   half-padded last dJ h-chunk) are held to their own scale.  The classic max-abs / max check is kept beside it at 3e-3
   (+ r_lat).

Config-4 width (T=750, U=200, V=5000: 20 vocabulary tiles, multi-chunk accumulate=1, the k=3 multi-item dW schedule)
is covered against the CPU reference chain both with default chunking (B=2, two chunks) and with three forced chunks.
"""
import numpy as np
import pytest
import torch

import tsasr_b200
from tsasr_b200 import _lib, ops
from oracle import image_decode as imd
from oracle.reference_chain import reference_joint_loss_fwd_bwd

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4       # north_star: per-utterance loss within 1e-4 relative
DLOGITS_ATOL = 1e-3    # north_star: dlogits within 1e-3 max-abs (fp32-accumulated) ...
DLOGITS_RTOL = 2.0 ** -8  # ... plus the rounding of the stored bf16 value
EXACT_RTOL = 1e-5      # fp64 redo of the GEMMs on the decoded images, relative to sum |terms| per element
K_SIGMA = 8.0          # elementwise statistical bound on the effect of bf16-rounded dlogits
R_SYS = 2e-4           # common-mode bound (approximate exp2 / log2, fp32 accumulation), relative to sum |terms|
# fp32 lattice floor: alpha / beta / L are fp32 numbers of magnitude |L| in BOTH implementations, and every dlogits row
# carries exp(alpha + beta - L).  tools/fp32_noise_floor.py measures the reference (torchaudio fp32) against float64:
# the relative error of a cell's gradient row reaches 0.2 * 2^-23 |L| sqrt(T+U) (2.6e-3 at configs[1] magnitudes,
# 6.9e-3 at configs[3]).  Two fp32 implementations may differ by twice that; C_LAT = 0.5 leaves a small margin.
C_LAT = 0.5
MAX_REL = 3e-3         # max |err| / max |ref| (round 1 used 1e-2), plus the fp32 lattice floor above


def _lattice_floor(costs, T, U):
    return C_LAT * 2.0 ** -23 * float(costs.abs().max()) * (T + U) ** 0.5


def _dev():
    return torch.device("cuda:0")


def _inputs(B, T, U, H, V, seed, ragged=True):
    gen = torch.Generator().manual_seed(seed)
    enc = (0.5 * torch.randn(B, T, H, generator=gen)).bfloat16()
    dec = (0.5 * torch.randn(B, U, H, generator=gen)).bfloat16()
    bound = 1.0 / (H ** 0.5)
    W = ((torch.rand(V, H, generator=gen) * 2 - 1) * bound).bfloat16()
    b = (torch.rand(V, generator=gen) * 2 - 1) * bound
    targets = torch.randint(1, V, (B, max(U - 1, 0)), generator=gen, dtype=torch.int32)
    ll = torch.full((B,), T, dtype=torch.int32)
    tl = torch.full((B,), U - 1, dtype=torch.int32)
    if ragged and B > 1:
        ll[1:] = torch.randint(max(1, T // 2), T + 1, (B - 1,), generator=gen, dtype=torch.int32)
        tl[1:] = torch.randint(0, U, (B - 1,), generator=gen, dtype=torch.int32)
    return enc, dec, W, b, targets, ll, tl


def _run_fused(enc, dec, W, b, targets, ll, tl, dcost, act, chunk, eps=None):
    d = _dev()
    e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
    costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, activation=act,
                                             reduction="none", max_chunk_cells=chunk, prune_log2_eps=eps)
    (costs * dcost.to(d)).sum().backward()
    torch.cuda.synchronize()
    return costs.detach().cpu(), {"d_enc": e.grad.cpu(), "d_dec": dc.grad.cpu(), "dW": w.grad.cpu(), "db": bb.grad.cpu()}


def _elementwise_report(got, ref, bounds, r_lat, dtype=torch.float64):
    """Per tensor: worst |got - ref| / (K_SIGMA sigma + (R_SYS + r_lat) sum|terms|) over all elements, and max/max."""
    rep = {}
    for k in ("d_enc", "d_dec", "dW", "db"):
        g, r = got[k].to(dtype), ref[k].to(dtype)
        err = (g - r).abs()
        if "sq_" + k in bounds:   # statistical bound on the bf16 rounding of dlogits
            rounding = K_SIGMA * bounds["sq_" + k].to(dtype).sqrt() * (2.0 ** -9 / 3 ** 0.5)
        else:                     # large shapes (no GEMM of squares): the worst case of that rounding, 2^-9 sum|terms|
            rounding = 2.0 ** -9 * bounds["abs_" + k].to(dtype)
        tol = rounding + (R_SYS + r_lat) * bounds["abs_" + k].to(dtype) + 1e-30
        rep[k] = (round((err / tol).max().item(), 3), float("%.2e" % (err.max().item() / max(r.abs().max().item(), 1e-30))))
    return rep


def _assert_elementwise(got, ref, bounds, tag, r_lat):
    rep = _elementwise_report(got, ref, bounds, r_lat)
    print(tag, "elementwise (worst err/tol, max/max):", rep)
    assert all(v[0] <= 1.0 for v in rep.values()), (tag, "elementwise bound exceeded", rep)
    assert all(v[1] < MAX_REL + r_lat for v in rep.values()), (tag, "max/max", rep)


def _assert_exact(got, exact, tag, n_cells=0):
    # fp32 accumulation (TMEM + split-K folds): the rounding error of a sum of n terms grows like sqrt(n) * 2^-24 of the
    # sum of |terms|; 1e-5 covers every contraction up to ~1e4 cells, longer ones (dW / db over a whole batch) get the sqrt
    rtol = EXACT_RTOL * max(1.0, (n_cells / 1.0e4) ** 0.5)
    rep = {k: round(((got[k].double() - exact[k]).abs() / (rtol * exact["abs_" + k] + 1e-30)).max().item(), 3)
           for k in ("d_enc", "d_dec", "dW", "db")}
    print(tag, f"fp64 GEMM on the decoded images (worst err / ({rtol:.1e} sum|terms|)):", rep)
    assert all(v <= 1.0 for v in rep.values()), (tag, "fp64 GEMM on the decoded images", rep)


SHAPES = [((2, 24, 9, 64, 40), "leaky_relu"), ((2, 16, 6, 64, 33), "tanh"), ((2, 40, 17, 640, 1000), "leaky_relu"),
          ((2, 30, 40, 320, 29), "relu"), ((3, 20, 1, 128, 50), "leaky_relu"), ((1, 9, 130, 384, 257), "identity"),
          ((3, 300, 80, 128, 500), "leaky_relu")]


@pytest.mark.parametrize("shape,act", SHAPES)
@pytest.mark.parametrize("eps", [0.0, -30.0], ids=["dense", "pruned"])
def test_backward_elementwise_vs_reference_and_exact_vs_images(shape, act, eps):
    B, T, U, H, V = shape
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=sum(shape) + 5)
    dcost = torch.linspace(0.5, 1.5, B)
    costs, got = _run_fused(enc, dec, W, b, targets, ll, tl, dcost, act, 1 << 40, eps)
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, act, 0.01, round_bf16=True, dcost=dcost,
                                       keep_intermediates=True)
    np.testing.assert_allclose(costs.numpy(), ref["costs"].numpy(), rtol=LOSS_RTOL)
    live = imd.live_cell_mask(B, T, U, ll, tl)
    # (3) end to end, elementwise
    bounds = imd.backward_from_operands(ref["dlogits"], ref["joint"], W.float(), enc.float(), dec.float(), act, 0.01, live)
    _assert_elementwise(got, ref, bounds, f"{shape} {act} eps={eps}", _lattice_floor(ref["costs"], T, U))
    # (1) dlogits images, element by element
    d = _dev()
    ws = ops.last_workspace(d)
    dY, J = imd.decode_images(ws, B, T, U, H, V)
    mask = live
    if eps < 0:
        off = int(_lib.load().tsasr_joint_bwd_stats_offset(B, T, U, H, V, 1 << 40))
        mask = imd.active_cell_mask(ws, off, B, T, U) & live
        pruned = live & ~mask
        if pruned.any():  # what pruning drops is negligible in the REFERENCE's gradient
            assert (ref["dlogits"].abs() * pruned[..., None]).max().item() < 1e-7
    m = mask[..., None]
    zero = torch.zeros(())
    dY, J = torch.where(m, dY, zero), torch.where(m, J, zero)  # tiles the kernels never wrote hold stale bytes (possibly NaN patterns)
    err = torch.where(m, (dY - ref["dlogits"]).abs() - DLOGITS_ATOL - (DLOGITS_RTOL + _lattice_floor(ref["costs"], T, U)) * ref["dlogits"].abs(), zero)
    assert err.max().item() <= 0.0, (shape, "dlogits images", torch.where(m, (dY - ref["dlogits"]).abs(), zero).max().item())
    assert torch.equal(J, torch.where(m, ref["joint"], zero)), "J operand images must be the bf16-rounded joint tensor exactly"
    # (2) the GEMMs / reductions on exactly those operands
    exact = imd.backward_from_operands(dY, J, W.float(), enc.float(), dec.float(), act, 0.01, mask)
    _assert_exact(got, exact, f"{shape} {act} eps={eps}", int(mask.sum()))


@pytest.mark.timeout(900)
def test_dlogits_images_ragged_config2_pair_with_pruning():
    """Two utterances at the full T, U, V, H of BASELINE configs[1], one of them short (T_b=263, 57 labels), tile
    pruning on: dlogits element by element against the reference's, and the fp64 redo of the GEMMs."""
    B, T, U, H, V = 2, 400, 100, 640, 1000
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=77)
    ll[1], tl[1] = 263, 57
    dcost = torch.tensor([1.0, 0.5])
    costs, got = _run_fused(enc, dec, W, b, targets, ll, tl, dcost, "leaky_relu", 1 << 40, -30.0)
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=True, dcost=dcost,
                                       keep_intermediates=True)
    np.testing.assert_allclose(costs.numpy(), ref["costs"].numpy(), rtol=LOSS_RTOL)
    d = _dev()
    ws = ops.last_workspace(d)
    dY, J = imd.decode_images(ws, B, T, U, H, V)
    live = imd.live_cell_mask(B, T, U, ll, tl)
    off = int(_lib.load().tsasr_joint_bwd_stats_offset(B, T, U, H, V, 1 << 40))
    mask = imd.active_cell_mask(ws, off, B, T, U) & live
    assert 0 < int(mask.sum()) < int(live.sum())  # pruning did skip tiles
    assert (ref["dlogits"].abs() * (live & ~mask)[..., None]).max().item() < 1e-7
    m = mask[..., None]
    zero = torch.zeros(())
    dY, J = torch.where(m, dY, zero), torch.where(m, J, zero)
    err = torch.where(m, (dY - ref["dlogits"]).abs() - DLOGITS_ATOL - (DLOGITS_RTOL + _lattice_floor(ref["costs"], T, U)) * ref["dlogits"].abs(), zero)
    assert err.max().item() <= 0.0, torch.where(m, (dY - ref["dlogits"]).abs(), zero).max().item()
    exact = imd.backward_from_operands(dY, J, W.float(), enc.float(), dec.float(), "leaky_relu", 0.01, mask, dtype=torch.float64,
                                       with_bounds=True)
    _assert_exact(got, exact, "config-2 pair", int(mask.sum()))
    bounds = imd.backward_from_operands(ref["dlogits"], ref["joint"], W.float(), enc.float(), dec.float(), "leaky_relu", 0.01, live,
                                        dtype=torch.float32)
    _assert_elementwise(got, ref, bounds, "config-2 pair", _lattice_floor(ref["costs"], T, U))


@pytest.mark.timeout(1500)
@pytest.mark.parametrize("B,chunk", [(1, 128 * 400), (2, 0)], ids=["B1-three-forced-chunks", "B2-default-two-chunks"])
def test_config4_width_fused_path_vs_reference_chain(B, chunk):
    """BASELINE configs[3] width: T=750, U=200, V=5000, H=640 (NT=20 vocabulary tiles, several backward chunks with
    accumulate=1, the multi-item dW schedule) against the reference chain on CPU (3 GB of fp32 logits per utterance)."""
    T, U, H, V = 750, 200, 640, 5000
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=40 + B, ragged=False)
    if B > 1:
        ll[1], tl[1] = 533, 140
    dcost = torch.linspace(0.5, 1.5, B)
    _lib.kernel_timing(True)  # the library's own launch record: how many backward chunks actually ran
    costs, got = _run_fused(enc, dec, W, b, targets, ll, tl, dcost, "leaky_relu", chunk)
    launched = _lib.kernel_timings()
    _lib.kernel_timing(False)
    assert launched["dw_gemm_kernel"][1] >= (3 if chunk else 2), launched  # accumulate=1 ran in the later chunks
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=True, dcost=dcost,
                                       keep_intermediates=True)
    np.testing.assert_allclose(costs.numpy(), ref["costs"].numpy(), rtol=LOSS_RTOL)
    live = imd.live_cell_mask(B, T, U, ll, tl)
    bounds = imd.backward_from_operands(ref["dlogits"], ref["joint"], W.float(), enc.float(), dec.float(), "leaky_relu", 0.01, live,
                                        dtype=torch.float32, with_sq=False)
    _assert_elementwise(got, ref, bounds, f"config-4 width B={B} chunk={chunk}", _lattice_floor(ref["costs"], T, U))
    for bi in range(B):  # exact zeros outside the utterance's rectangle
        assert not got["d_enc"][bi, int(ll[bi]):].any() and not got["d_dec"][bi, int(tl[bi]) + 1:].any()
