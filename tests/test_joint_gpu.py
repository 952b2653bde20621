"""GPU parity tests of the fused joint + RNN-T loss path (tcgen05 kernels) through the C ABI.

Oracle: the reference chain restated on CPU (oracle/reference_chain.py: Transducer_joint "sum" ->
Linear -> torchaudio rnnt_loss), fed the SAME bf16-rounded operands the kernels see.
Tolerances (north_star): per-utterance loss 1e-4 relative; gradients: dlogits-level 1e-3 max-abs."""
import numpy as np
import pytest
import torch

import tsasr_b200
from tsasr_b200 import _lib, ops
from oracle.reference_chain import reference_joint_logits, reference_joint_loss_fwd_bwd

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4


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


@pytest.mark.parametrize("shape,act", [((2, 24, 9, 64, 40), "leaky_relu"), ((1, 16, 8, 128, 29), "tanh"),
                                       ((2, 40, 20, 640, 1000), "leaky_relu"), ((3, 17, 5, 256, 300), "relu"),
                                       ((1, 33, 100, 192, 1031), "identity")])
def test_tcgen05_logits_match_reference_gemm(shape, act):
    """The TMA/tcgen05/TMEM mainloop alone: recomputed logits vs a torch fp32 GEMM on the same bf16 operands."""
    B, T, U, H, V = shape
    enc, dec, W, b, *_ = _inputs(B, T, U, H, V, seed=sum(shape))
    d = _dev()
    got = ops.joint_debug_logits(enc.to(d), dec.to(d), W.to(d), b.to(d), _lib.ACT_CODES[act], 0.01).cpu()
    ref = reference_joint_logits(enc.float(), dec.float(), W.float(), b, act, 0.01, round_bf16=True)
    err = (got - ref).abs().max().item()
    assert err < 2e-3, f"max |logits - ref| = {err}"


@pytest.mark.parametrize("shape,act", [((2, 24, 9, 64, 40), "leaky_relu"), ((2, 16, 6, 64, 33), "tanh"),
                                       ((4, 50, 17, 640, 1000), "leaky_relu"), ((2, 30, 40, 320, 29), "relu"),
                                       ((3, 20, 1, 128, 50), "leaky_relu"), ((2, 9, 130, 64, 257), "leaky_relu")])
def test_fused_forward_costs_vs_reference_chain(shape, act):
    B, T, U, H, V = shape
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=sum(shape) + 1)
    d = _dev()
    costs = tsasr_b200.fused_joint_rnnt_loss(enc.to(d), dec.to(d), W.to(d), b.to(d), targets.to(d), ll.to(d), tl.to(d),
                                             blank=0, activation=act, reduction="none")
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, act, 0.01, round_bf16=True)
    np.testing.assert_allclose(costs.cpu().numpy(), ref["costs"].numpy(), rtol=LOSS_RTOL)


def test_fused_forward_lattice_vs_compat_path():
    """The fused epilogue's {lp_blank, lp_emit, logZ} equal the compat kernels' on the debug logits."""
    B, T, U, H, V = 2, 35, 11, 128, 500
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=99)
    d = _dev()
    args = (enc.to(d), dec.to(d), W.to(d), b.to(d))
    lat_f, logz_f = ops.joint_fwd(*args, targets.to(d), ll.to(d), tl.to(d), 0, 0, 0.01)
    logits = ops.joint_debug_logits(*args, 0, 0.01)
    lat_c, den_c = ops.logits_to_lattice(logits, targets.to(d), ll.to(d), tl.to(d), 0)
    lf, lc = ops.unskew(lat_f, B, T, U), ops.unskew(lat_c, B, T, U)
    zf, zc = ops.unskew(logz_f, B, T, U), ops.unskew(den_c, B, T, U)
    for bi in range(B):
        Tb, Ub = int(ll[bi]), int(tl[bi]) + 1
        torch.testing.assert_close(zf[bi, :Tb, :Ub], zc[bi, :Tb, :Ub], rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(lf[bi, :Tb, :Ub, 0], lc[bi, :Tb, :Ub, 0], rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(lf[bi, :Tb, :Ub - 1, 1], lc[bi, :Tb, :Ub - 1, 1], rtol=1e-5, atol=2e-5)
