"""Full-size checks of the fused path at BASELINE.json's configurations (configs[1] and configs[3]).

The CPU oracle cannot run B*T*U*V = 6.4e8 .. 6e9 logits in test time, so these tests use
  * one FULL-SIZE utterance (B=1, T=400, U=100, V=1000, H=640) against the reference chain on CPU, and
  * size-independent properties of the whole batch: alpha/beta agreement, exact zeros outside T_b x U_b,
    sum_v db[v] = 0 (every dlogits row sums to zero), exact linearity in the upstream gradient,
    invariance of an utterance's results to what else is in the batch, and bounded memory (no 4-D tensor).
All calls go through the C ABI (ops / functional)."""
import numpy as np
import pytest
import torch

import tsasr_b200
from tsasr_b200 import _lib, ops
from oracle.reference_chain import reference_joint_loss_fwd_bwd

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _inputs(B, T, U, H, V, seed, ragged):
    g = torch.Generator().manual_seed(seed)
    enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16()
    dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16()
    bound = 1.0 / H ** 0.5
    W = ((torch.rand(V, H, generator=g) * 2 - 1) * bound).bfloat16()
    b = (torch.rand(V, generator=g) * 2 - 1) * bound
    targets = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32)
    ll = torch.full((B,), T, dtype=torch.int32)
    tl = torch.full((B,), U - 1, dtype=torch.int32)
    if ragged and B > 1:  # T_b in [0.6 T, T], label count in [0.4 U, U-1], one row at the maximum each
        ll[1:] = torch.randint(int(0.6 * T), T + 1, (B - 1,), generator=g, dtype=torch.int32)
        tl[1:] = torch.randint(int(0.4 * U), U, (B - 1,), generator=g, dtype=torch.int32)
    return enc, dec, W, b, targets, ll, tl


def _run(enc, dec, W, b, targets, ll, tl, dcost, act="leaky_relu", max_chunk_cells=0):
    d = _dev()
    e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
    costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, activation=act,
                                             reduction="none", max_chunk_cells=max_chunk_cells)
    (costs * dcost.to(d)).sum().backward()
    torch.cuda.synchronize()
    return costs.detach(), e.grad, dc.grad, w.grad, bb.grad


def _rel(got, ref):
    return (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)


@pytest.mark.timeout(600)
def test_config2_single_full_size_utterance_vs_reference_chain():
    """B=1 at the full T, U, V, H of configs[1]: per-utterance loss 1e-4 relative, operand gradients 3e-3 of max (+ fp32 lattice floor)."""
    B, T, U, H, V = 1, 400, 100, 640, 1000
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=11, ragged=False)
    dcost = torch.ones(B)
    costs, d_enc, d_dec, dW, db = _run(enc, dec, W, b, targets, ll, tl, dcost)
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=True, dcost=dcost)
    np.testing.assert_allclose(costs.cpu().numpy(), ref["costs"].numpy(), rtol=1e-4)
    errs = {"d_enc": _rel(d_enc.cpu(), ref["d_enc"]), "d_dec": _rel(d_dec.cpu(), ref["d_dec"]),
            "dW": _rel(dW.cpu(), ref["dW"]), "db": _rel(db.cpu(), ref["db"])}
    # 3e-3 of the largest entry + the floor an fp32 lattice of this size imposes on any implementation, the reference
    # included (tools/fp32_noise_floor.py; elementwise bounds: test_parity_tight_gpu.py)
    floor = 0.5 * 2.0 ** -23 * float(ref["costs"].abs().max()) * (T + U) ** 0.5
    assert all(v < 3e-3 + floor for v in errs.values()), (errs, floor)


def _properties(B, T, U, H, V, seed, max_chunk_cells=0):
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=seed, ragged=True)
    d = _dev()
    dcost = torch.linspace(0.5, 1.5, B) / B
    torch.cuda.reset_peak_memory_stats(d)
    base = torch.cuda.memory_allocated(d)
    costs, d_enc, d_dec, dW, db = _run(enc, dec, W, b, targets, ll, tl, dcost, max_chunk_cells=max_chunk_cells)
    peak = torch.cuda.max_memory_allocated(d) - base

    # (1) the two DP directions agree on log P(y|x) for every utterance (what the reference asserts implicitly:
    #     torchaudio takes the cost from beta(0,0), the Numba kernels from alpha(T-1,U-1))
    lat2, logz = ops.joint_fwd(enc.to(d), dec.to(d), W.to(d), b.to(d), targets.to(d), ll.to(d), tl.to(d), 0,
                               _lib.ACT_CODES["leaky_relu"], 0.01)
    _, _, cost2, ll_a, ll_b = ops.alpha_beta(lat2, ll.to(d), tl.to(d), B, T, U)
    np.testing.assert_allclose(ll_a.cpu().numpy(), ll_b.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(costs.cpu().numpy(), cost2.cpu().numpy(), rtol=1e-6)
    assert torch.isfinite(costs).all() and (costs > 0).all()

    # (2) exact zeros outside every utterance's T_b x U_b rectangle
    for bi in range(B):
        assert not d_enc[bi, int(ll[bi]):].any() and not d_dec[bi, int(tl[bi]) + 1:].any()

    # (3) every dlogits row sums to zero (softmax gradient), hence sum_v db[v] = 0 up to bf16 rounding of dY
    assert abs(db.sum().item()) <= 2e-3 * db.abs().sum().item(), (db.sum().item(), db.abs().sum().item())

    # (4) exact linearity in the upstream gradient: doubling dcost doubles every gradient bit for bit
    _, d_enc2, d_dec2, dW2, db2 = _run(enc, dec, W, b, targets, ll, tl, 2 * dcost, max_chunk_cells=max_chunk_cells)
    for g1, g2 in ((d_enc, d_enc2), (d_dec, d_dec2), (dW, dW2), (db, db2)):
        assert torch.equal(2 * g1, g2)

    # (5) an utterance's loss and operand gradients do not depend on the rest of the batch
    #     (utterance 0 has the maximal lengths, so the sub-batch keeps the padded T and U)
    idx = torch.tensor([0, B - 1])
    costs_s, d_enc_s, d_dec_s, _, _ = _run(enc[idx], dec[idx], W, b, targets[idx], ll[idx], tl[idx], dcost[idx],
                                           max_chunk_cells=max_chunk_cells)
    np.testing.assert_allclose(costs_s.cpu().numpy(), costs[idx.to(costs.device)].cpu().numpy(), rtol=1e-6)
    assert _rel(d_enc_s, d_enc[idx.to(d)]) < 1e-5 and _rel(d_dec_s, d_dec[idx.to(d)]) < 1e-5
    return peak


@pytest.mark.timeout(600)
def test_config2_full_batch_properties():
    """configs[1]: B=16, T=400, U=100, V=1000, H=640, ragged lengths."""
    peak = _properties(16, 400, 100, 640, 1000, seed=3)
    logits_bytes = 16 * 400 * 100 * 1000 * 4
    # operand images (bf16 dlogits + J, one chunk) + lattice arrays; the reference holds >= 3 fp32 [B,T,U,V] tensors
    assert peak < 1.2 * logits_bytes, peak


@pytest.mark.timeout(900)
def test_config4_long_mixture_properties_and_memory():
    """configs[3]: B=8, T=750, U=200, V=5000 -- the 4-D logits (24 GB fp32) must never exist: the backward runs in
    bounded chunks of operand images (default 3 GiB each)."""
    peak = _properties(8, 750, 200, 640, 5000, seed=4)
    logits_bytes = 8 * 750 * 200 * 5000 * 4
    assert peak < 0.25 * logits_bytes, (peak, logits_bytes)
