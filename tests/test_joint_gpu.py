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


# ---------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------
def _rel_err(got, ref):
    return (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)


GRAD_REL = 3e-3  # max |err| / max |ref| of the operand gradients (bf16 dlogits / J operands, fp32 accumulation); the elementwise
                 # bounds and the fp64 redo of the GEMMs on the decoded operand images live in test_parity_tight_gpu.py


def _choose_tile_log2(T, U):
    best, best_l = -1, 4
    for l in range(3, 6):
        tT, tU = 1 << l, 128 >> l
        padded = ((T + tT - 1) // tT) * tT * ((U + tU - 1) // tU) * tU
        if best < 0 or padded < best or (padded == best and l == 4):
            best, best_l = padded, l
    return best_l


def _decode_dy_images(ws, B, T, U, V):
    """Rebuild dense dlogits [B,T,U,V] from the bf16 SWIZZLE_128B operand images in the workspace
    (single-chunk runs only) -- the test-only view of the never-materialised dlogits."""
    l = _choose_tile_log2(T, U)
    tT, tU = 1 << l, 128 >> l
    nTt, nTu = (T + tT - 1) // tT, (U + tU - 1) // tU
    NT4 = 4 * ((V + 255) // 256)
    n_tiles = B * nTt * nTu
    base = (-ws.data_ptr()) % 1024
    raw = ws[base: base + n_tiles * NT4 * 16384].view(torch.bfloat16).view(n_tiles, NT4, 128, 64).float().cpu()
    # un-swizzle: 16-byte chunk c of row r is stored at chunk c ^ (r & 7)
    r = torch.arange(128)[:, None]
    c = torch.arange(8)[None, :]
    src_chunk = (c ^ (r & 7))  # [128, 8]
    idx = (src_chunk[:, :, None] * 8 + torch.arange(8)[None, None, :]).reshape(128, 64)
    img = torch.gather(raw, 3, idx[None, None].expand(n_tiles, NT4, 128, 64))
    out = torch.zeros(B, T, U, V)
    for tile in range(n_tiles):
        b, rem = divmod(tile, nTt * nTu)
        tt, tu = divmod(rem, nTu)
        rows = img[tile].permute(1, 0, 2).reshape(128, NT4 * 64)[:, :V]  # [128 rows, V]
        for row in range(128):
            ti, ui = row % tT, row // tT
            t, u = tt * tT + ti, tu * tU + ui
            if t < T and u < U:
                out[b, t, u] = rows[row]
    return out


@pytest.mark.parametrize("shape,act", [((2, 24, 9, 64, 40), "leaky_relu"), ((2, 16, 6, 64, 33), "tanh"),
                                       ((2, 40, 17, 640, 1000), "leaky_relu"), ((2, 30, 40, 320, 29), "relu"),
                                       ((3, 20, 1, 128, 50), "leaky_relu"), ((1, 9, 130, 384, 257), "leaky_relu"),
                                       # H not a multiple of 64: zero-padded to whole k-blocks by the fused op
                                       ((2, 18, 7, 100, 90), "tanh"), ((2, 12, 5, 40, 30), "leaky_relu"),
                                       ((1, 20, 9, 600, 300), "relu")])
def test_fused_backward_vs_reference_chain(shape, act):
    B, T, U, H, V = shape
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=sum(shape) + 2)
    d = _dev()
    dcost = torch.linspace(0.5, 1.5, B)
    e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
    costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, activation=act,
                                             reduction="none", max_chunk_cells=1 << 30)
    (costs * dcost.to(d)).sum().backward()
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, act, 0.01, round_bf16=True, dcost=dcost)
    np.testing.assert_allclose(costs.detach().cpu().numpy(), ref["costs"].numpy(), rtol=LOSS_RTOL)
    errs = {k: _rel_err(g.grad.cpu(), ref[k]) for k, g in (("d_enc", e), ("d_dec", dc), ("dW", w), ("db", bb))}
    assert all(v < GRAD_REL for v in errs.values()), errs
    # padding rows of d_enc / d_dec are exactly zero
    for bi in range(B):
        assert not e.grad[bi, int(ll[bi]):].any() and not dc.grad[bi, int(tl[bi]) + 1:].any()


def test_fused_dlogits_images_vs_oracle():
    """Test-only view of the tile-wise dlogits (bf16 operand images): 1e-3 max-abs (north_star) plus the
    bf16 rounding of the value itself (2^-8 relative)."""
    from oracle import rnnt_numpy as rn

    B, T, U, H, V = 2, 20, 7, 64, 90
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=5)
    d = _dev()
    e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
    costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0,
                                             reduction="none", max_chunk_cells=1 << 30, prune_log2_eps=0.0)  # every tile written
    costs.sum().backward()
    torch.cuda.synchronize()
    got = _decode_dy_images(ops.last_workspace(d), B, T, U, V).numpy()
    _, logits = rn.joint_logits(enc.float().numpy(), dec.float().numpy(), W.float().numpy(), b.numpy(), "leaky_relu")
    _, want = rn.rnnt_torchaudio(logits, targets.numpy(), ll.numpy(), tl.numpy(), 0)
    for bi in range(B):  # tiles entirely outside the T_b x U_b rectangle are never written (nor read)
        Tb, Ub = int(ll[bi]), int(tl[bi]) + 1
        g_, w_ = got[bi, :Tb, :Ub], want[bi, :Tb, :Ub]
        assert np.all(np.abs(g_ - w_) <= 1e-3 + np.abs(w_) * 2.0 ** -8), np.abs(g_ - w_).max()


def test_fused_backward_chunked_equals_single_chunk():
    B, T, U, H, V = 3, 48, 20, 128, 200
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=21)
    d = _dev()
    grads = []
    for chunk in (1 << 30, 128 * 7):
        e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
        loss = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0,
                                                reduction="mean", max_chunk_cells=chunk)
        loss.backward()
        grads.append([g.grad.clone() for g in (e, dc, w, bb)])
    for a, c in zip(*grads):
        assert _rel_err(c, a) < 1e-5  # same arithmetic, different fp32 summation grouping


@pytest.mark.parametrize("name", ["joint_leaky", "joint_tanh", "joint_relu"])
def test_drop_in_modules_vs_golden(golden, name):
    """Transducer_joint -> stock Linear-style head -> transducer_loss, exactly the recipe's three
    call sites (train_librispeechmix_scratch.py:132,135,158), against vectors produced by the reference."""
    g = golden(name)
    d = _dev()
    act = {"leaky_relu": torch.nn.LeakyReLU, "tanh": torch.nn.Tanh, "relu": torch.nn.ReLU}[str(g["act"])]
    enc = torch.tensor(g["enc"], device=d, requires_grad=True)
    dec = torch.tensor(g["dec"], device=d, requires_grad=True)
    V, H = g["W"].shape
    head = torch.nn.Linear(H, V).to(d)
    with torch.no_grad():
        head.weight.copy_(torch.tensor(g["W"]))
        head.bias.copy_(torch.tensor(g["b"]))
    joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=act)
    joint = joiner(enc[..., None, :], dec[:, None, ...])
    assert isinstance(joint, tsasr_b200.JointHandle)
    logits = head(joint)
    assert isinstance(logits, tsasr_b200.JointHandle) and logits.shape == (*joint.shape[:3], V)
    loss = tsasr_b200.transducer_loss(logits, torch.tensor(g["targets"], device=d).long(),
                                      torch.tensor(g["input_rel"], device=d), torch.tensor(g["target_rel"], device=d),
                                      blank_index=0, reduction=str(g["reduction"]), use_torchaudio=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=LOSS_RTOL)
    for got, key in ((enc.grad, "d_enc"), (dec.grad, "d_dec"), (head.weight.grad, "dW"), (head.bias.grad, "db")):
        assert _rel_err(got.cpu(), torch.tensor(g[key])) < GRAD_REL, key


@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
@pytest.mark.parametrize("entry", ["transducer_loss", "TransducerLoss"])
def test_numba_semantics_through_the_handle_stay_fused(entry, reduction):
    """``use_torchaudio=False`` / ``TransducerLoss`` on a deferred handle: value reduce_b(-log P_b / T_b) and the
    un-normalised gradient of the reference's Numba branch (SB/nnet/loss/transducer_loss.py:104-106,280-293), computed
    by the fused kernels.  Checked against (1) the materialised Numba-semantics path (itself pinned on golden vectors
    produced by the reference's own kernels) and (2) the CPU reference chain with d cost_b = 1."""
    B, T, U, H, V = 3, 28, 11, 128, 60
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=4242)
    d = _dev()
    joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)

    def run(fused):
        e = enc.float().to(d).requires_grad_()
        dc = dec.float().to(d).requires_grad_()
        head = torch.nn.Linear(H, V).to(d)
        with torch.no_grad():
            head.weight.copy_(W.float())
            head.bias.copy_(b)
        logits = head(joiner(e[..., None, :], dc[:, None, ...]))
        assert isinstance(logits, tsasr_b200.JointHandle)
        if not fused:
            logits = logits.materialize()
        if entry == "TransducerLoss":
            loss = tsasr_b200.TransducerLoss(blank=0, reduction=reduction)(logits, targets.to(d), ll.to(d), tl.to(d))
        else:
            loss = tsasr_b200.transducer_loss(logits, targets.to(d), (ll.float() / T).to(d), (tl.float() / (U - 1)).to(d),
                                              blank_index=0, reduction=reduction, use_torchaudio=False)
        gout = torch.arange(1, B + 1, dtype=torch.float32, device=d) if reduction == "none" else None
        loss.backward(gout)
        return loss.detach().cpu(), [x.grad.cpu() for x in (e, dc, head.weight, head.bias)]

    if entry == "transducer_loss":  # relative lengths must convert back to the same integers
        assert torch.equal(((ll.float() / T) * T).round().int(), ll) and torch.equal(((tl.float() / (U - 1)) * (U - 1)).round().int(), tl)
    loss_f, grads_f = run(True)
    loss_m, grads_m = run(False)
    np.testing.assert_allclose(loss_f.numpy(), loss_m.numpy(), rtol=LOSS_RTOL)
    for a, c in zip(grads_f, grads_m):
        assert _rel_err(a, c) < GRAD_REL
    dcost = torch.arange(1, B + 1, dtype=torch.float32) if reduction == "none" else torch.ones(B)
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=True, dcost=dcost)
    per_utt = ref["costs"] / ll.float()
    want = {"mean": per_utt.mean(), "sum": per_utt.sum(), "none": per_utt}[reduction]
    np.testing.assert_allclose(loss_f.numpy(), want.numpy(), rtol=LOSS_RTOL)
    for got, key in zip(grads_f, ("d_enc", "d_dec", "dW", "db")):
        assert _rel_err(got, ref[key]) < GRAD_REL, key


def test_fused_length_precondition_errors_are_raised_after_safe_launch():
    """The fused path validates lengths on a side stream AFTER queueing its kernels (they clamp every length to the
    padded lattice): the same RuntimeErrors as torchaudio must come out and the device must stay healthy."""
    B, T, U, H, V = 2, 24, 9, 64, 40
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=77, ragged=False)
    d = _dev()
    args = [x.to(d) for x in (enc.float(), dec.float(), W.float(), b, targets)]

    def call(ll_, tl_):
        return tsasr_b200.fused_joint_rnnt_loss(*args, ll_.to(d), tl_.to(d), blank=0, reduction="none")

    with pytest.raises(RuntimeError, match="input length mismatch"):
        call(torch.tensor([T + 500, T], dtype=torch.int32), tl)
    with pytest.raises(RuntimeError, match="input length mismatch"):
        call(torch.tensor([T - 1, T - 2], dtype=torch.int32), tl)
    with pytest.raises(RuntimeError, match="output length mismatch"):
        call(ll, torch.tensor([U + 300, U - 1], dtype=torch.int32))
    with pytest.raises(RuntimeError, match="logit_lengths must be >= 1"):
        call(torch.tensor([T, 0], dtype=torch.int32), tl)
    with pytest.raises(RuntimeError, match="logit_lengths must be >= 1"):
        call(ll, torch.tensor([U - 1, -3], dtype=torch.int32))
    torch.cuda.synchronize()
    costs = call(ll, tl)  # a valid call afterwards: finite, matches the reference
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=True)
    np.testing.assert_allclose(costs.cpu().numpy(), ref["costs"].numpy(), rtol=LOSS_RTOL)


def test_relative_length_conversion_kernel_is_bit_exact():
    """tsasr_prepare_lengths vs the reference expression of SB/nnet/losses.py:58-59 on half-way and random values:
    integer results must be identical (round-half-to-even of the fp32 product), statistics included."""
    from tsasr_b200.functional import _prepare_lengths

    d = _dev()
    g = torch.Generator().manual_seed(0)
    for T, n_tg in ((400, 99), (750, 199), (37, 6), (1, 1), (2, 7)):
        halves = (torch.arange(0, 2 * T + 1, dtype=torch.float32) / (2.0 * T))          # k / 2T: exact .5 products
        rel_l = torch.cat([halves, torch.rand(3000, generator=g), torch.tensor([1.0, 0.0, 1e-8, 0.99999994])])
        rel_t = torch.rand(rel_l.shape[0], generator=g)
        rel_t[: min(len(halves), 2 * n_tg + 1)] = (torch.arange(0, 2 * n_tg + 1, dtype=torch.float32) / (2.0 * n_tg))[: len(halves)]
        want_l = (rel_l * T).round().int()
        want_t = (rel_t * n_tg).round().int()
        ll, tl, stats, _ = _prepare_lengths(rel_l.to(d), rel_t.to(d), T, n_tg, relative=True)
        torch.cuda.synchronize()
        assert torch.equal(ll.cpu(), want_l) and torch.equal(tl.cpu(), want_t)
        assert stats.cpu().tolist() == [int(want_l.max()), int(want_t.max()), int(want_l.min()), int(want_t.min())]
        ll2, tl2, stats2, _ = _prepare_lengths(want_l.to(d), want_t.to(d), T, n_tg, relative=False)
        assert torch.equal(ll2.cpu(), want_l) and stats2.cpu().tolist() == stats.cpu().tolist()


def test_single_call_forward_converts_lengths_and_targets_bit_exactly():
    """tsasr_joint_loss_fwd (the whole forward behind one C call): its preparation kernel must produce exactly the
    integers of SB/nnet/losses.py:58-59 (fp32 product, round-half-to-even, int32) and of ``targets.int()`` (:74), on
    half-way cases, for thousands of utterances at once."""
    from tsasr_b200.functional import FusedJointRnnt

    d = _dev()
    g = torch.Generator().manual_seed(1)
    for T, U in ((37, 7), (400, 3)):
        n_tg = U - 1
        halves = torch.arange(1, 2 * T + 1, dtype=torch.float32) / (2.0 * T)           # k / 2T: exact .5 products
        rel_l = torch.cat([halves, torch.rand(1500, generator=g).clamp(min=0.02), torch.tensor([1.0, 0.99999994])])
        B = rel_l.shape[0]
        rel_t = torch.rand(B, generator=g)
        rel_t[: 2 * n_tg + 1] = torch.arange(0, 2 * n_tg + 1, dtype=torch.float32) / (2.0 * n_tg)
        want_l, want_t = (rel_l * T).round().int(), (rel_t * n_tg).round().int()
        H, V = 64, 5
        enc = torch.zeros(B, T, H, device=d)
        dec = torch.zeros(B, U, H, device=d)
        W = torch.zeros(V, H, device=d)
        bias = torch.zeros(V, device=d)
        targets = torch.randint(1, V, (B, n_tg), generator=g)                           # int64, as the recipe's tokens
        cost, ll, tl = FusedJointRnnt.apply(enc, dec, W, bias, targets.to(d), rel_l.to(d), rel_t.to(d), True, 0, 0, 0.01, 0, None, -1.0,
                                            False)
        torch.cuda.synchronize()
        assert torch.equal(ll.cpu(), want_l) and torch.equal(tl.cpu(), want_t)
        assert torch.isfinite(cost).all()
        # uniform logits: -log P is the number of lattice paths times V^-(steps) -- compare one utterance with the closed form
        b0 = int(torch.argmax(want_l.long() * 1000 + want_t.long()))
        Tb, Lb = int(want_l[b0]), int(want_t[b0])
        import math
        want = (Tb + Lb) * math.log(V) - math.log(math.comb(Tb - 1 + Lb, Lb))
        assert abs(cost[b0].item() - want) < 1e-4 * want


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_precision_operands_need_no_cast_and_return_their_dtype(dtype):
    """SURVEY 8f N1 (projections as producers): when encoder_proj / decoder_proj already emit bf16 (autocast), enc_out and
    dec_out enter the fused op as they are and their gradients come back in the same dtype."""
    B, T, U, H, V = 2, 24, 9, 64, 40
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=3)
    d = _dev()
    e, dc = (x.to(d).to(dtype).requires_grad_() for x in (enc, dec))
    w, bb = W.to(d).float().requires_grad_(), b.to(d).requires_grad_()
    costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, reduction="none")
    costs.sum().backward()
    assert e.grad.dtype == dtype and dc.grad.dtype == dtype and w.grad.dtype == torch.float32
    ref = reference_joint_loss_fwd_bwd(e.detach().float().cpu(), dc.detach().float().cpu(), W, b, targets, ll, tl, 0, "leaky_relu",
                                       0.01, round_bf16=True)
    tol = LOSS_RTOL if dtype == torch.bfloat16 else 2e-3  # fp16 inputs are re-rounded to bf16 operands
    np.testing.assert_allclose(costs.detach().cpu().numpy(), ref["costs"].numpy(), rtol=tol)
    assert _rel_err(e.grad.float().cpu(), ref["d_enc"]) < 2e-2 and _rel_err(w.grad.cpu(), ref["dW"]) < 2e-2


def test_backward_tile_pruning_is_invisible_in_fp32():
    """Tiles whose alignment posterior stays below 2^-30 are skipped by the backward: the gradients must agree with the
    unpruned run to fp32 rounding, while a substantial share of the tiles is actually skipped."""
    B, T, U, H, V = 3, 300, 80, 128, 500
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=9)
    d = _dev()
    outs = []
    for eps in (0.0, -30.0):
        e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
        costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, reduction="none",
                                                 prune_log2_eps=eps)
        costs.sum().backward()
        torch.cuda.synchronize()
        outs.append(([g.grad.clone() for g in (e, dc, w, bb)], ops.last_backward_tile_stats(d)))
    (dense, st0), (pruned, st1) = outs
    assert st0 == (0, 0)                                   # switched off: no activity pass at all
    assert 0 < st1[0] < 0.85 * st1[1], st1                 # the far corners of the lattice are skipped
    for a, c in zip(dense, pruned):
        assert _rel_err(c, a) < 1e-5


@pytest.mark.parametrize("V,blank", [(300, -1), (300, 137), (1000, 255), (1000, 256), (1000, 999), (40, 16), (257, 256)])
def test_fused_path_blank_anywhere_in_the_vocabulary(V, blank):
    """The blank column may sit in any 16-column chunk of any vocabulary tile (torchaudio's default is -1 = V-1): the
    epilogue finds it through the per-tile special-chunk mask.  Against the CPU reference chain (torchaudio rnnt_loss)."""
    B, T, U, H = 3, 21, 12, 128
    enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=7 * V + blank)
    bl = blank % V
    targets = torch.where(targets == bl, torch.full_like(targets, (bl + 1) % V), targets)  # labels never equal the blank
    d = _dev()
    e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
    costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=blank, reduction="none")
    dcost = torch.linspace(0.5, 1.5, B)
    (costs * dcost.to(d)).sum().backward()
    ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, bl, "leaky_relu", 0.01, round_bf16=True, dcost=dcost)
    np.testing.assert_allclose(costs.detach().cpu().numpy(), ref["costs"].numpy(), rtol=LOSS_RTOL)
    for got, key in ((e.grad, "d_enc"), (dc.grad, "d_dec"), (w.grad, "dW"), (bb.grad, "db")):
        assert _rel_err(got.cpu(), ref[key]) < GRAD_REL, key


def test_fused_path_fuzz_vs_compat_path():
    """Random shapes, ragged lengths, activations, chunk sizes: the fused tcgen05 path (tile pruning on) against the
    compat kernels (materialised logits built by torch from the same bf16-rounded operands; themselves pinned on the
    C oracle and the golden vectors in test_lattice_gpu.py).  Loss 1e-4 relative, gradients 3e-3 of their largest entry (same DP kernel on both sides: no fp32 lattice floor between them)."""
    from oracle.reference_chain import reference_joint_logits

    d = _dev()
    rng = np.random.default_rng(2024)
    acts = ["leaky_relu", "relu", "tanh", "identity"]
    for case in range(24):
        B = int(rng.integers(1, 5))
        T = int(rng.integers(1, 90))
        U = int(rng.integers(1, 45))
        H = 64 * int(rng.integers(1, 11))
        V = int(rng.integers(2, 1100))
        act = acts[case % 4]
        enc, dec, W, b, targets, ll, tl = _inputs(B, T, U, H, V, seed=1000 + case)
        dcost = torch.tensor(rng.uniform(0.2, 2.0, B), dtype=torch.float32)
        chunk = 0 if case % 3 else 128 * int(rng.integers(1, 9))
        e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
        costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, activation=act,
                                                 reduction="none", max_chunk_cells=chunk)
        (costs * dcost.to(d)).sum().backward()
        e2, dc2, w2, bb2 = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
        logits = reference_joint_logits(e2, dc2, w2, bb2, act, 0.01, round_bf16=True)
        costs2 = tsasr_b200.rnnt_loss(logits.contiguous(), targets.to(d), ll.to(d), tl.to(d), blank=0, reduction="none")
        (costs2 * dcost.to(d)).sum().backward()
        tag = f"case {case}: B={B} T={T} U={U} H={H} V={V} act={act} chunk={chunk}"
        np.testing.assert_allclose(costs.detach().cpu().numpy(), costs2.detach().cpu().numpy(), rtol=LOSS_RTOL, err_msg=tag)
        for got, ref, name in ((e.grad, e2.grad, "d_enc"), (dc.grad, dc2.grad, "d_dec"), (w.grad, w2.grad, "dW"), (bb.grad, bb2.grad, "db")):
            assert _rel_err(got, ref) < GRAD_REL, (tag, name, _rel_err(got, ref))


def test_concat_joint_with_linear_joint_network_is_fused_as_a_sum_of_projections():
    """``Transducer_joint(joint="concat", joint_network=Linear)`` (SB/nnet/transducer/transducer_joint.py:76-93): the
    drop-in applies the two halves of the Linear to the encoder / predictor rows and runs the fused "sum" path; the
    reference expands, concatenates and runs the Linear over all B*T*U rows.
    (1) The rewrite is exact: bit-identical to the fused "sum" joint fed with the two projections computed by hand.
    (2) Against the reference's eager math (fp32 throughout): loss within 3e-4; gradients in relative L2 -- the projected
        operands are not bf16-representable here, so wherever the rounding flips the sign of a pre-activation LeakyReLU's
        derivative jumps between 1 and 0.01 for that ELEMENT (the price of the bf16 joint BASELINE.json specifies), which
        a max-norm over a 9-term sum shows at the percent level while the L2 norm stays small."""
    d = _dev()
    g = torch.Generator().manual_seed(12)
    B, T, U, He, Hd, Hj, V = 3, 33, 9, 48, 80, 128, 60
    enc = (0.5 * torch.randn(B, T, He, generator=g)).to(d)
    dec = (0.5 * torch.randn(B, U, Hd, generator=g)).to(d)
    targets = torch.randint(1, V, (B, U - 1), generator=g).to(d)
    il = torch.tensor([1.0, 0.7, 0.9], device=d)
    tl = torch.tensor([1.0, 0.5, 0.25], device=d)

    def run(mode):
        torch.manual_seed(3)
        jn = torch.nn.Linear(He + Hd, Hj).to(d)
        head = torch.nn.Linear(Hj, V).to(d)
        joiner = tsasr_b200.Transducer_joint(joint_network=jn, joint="concat", nonlinearity=torch.nn.LeakyReLU)
        e, dc = enc.clone().requires_grad_(), dec.clone().requires_grad_()
        if mode == "fused":
            logits = head(joiner(e[..., None, :], dc[:, None, ...]))
            assert isinstance(logits, tsasr_b200.JointHandle) and logits.has_head
        elif mode == "by hand":  # the same two projections, then the fused "sum" joint
            summer = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
            ep = torch.nn.functional.linear(e[..., None, :], jn.weight[:, :He], jn.bias)
            dp = torch.nn.functional.linear(dc[:, None, ...], jn.weight[:, He:])
            logits = head(summer(ep, dp))
            assert isinstance(logits, tsasr_b200.JointHandle) and logits.has_head
        else:                    # the reference's expand + cat + Linear over all B*T*U rows
            logits = head(joiner._eager(e[..., None, :], dc[:, None, ...]))
            assert logits.shape == (B, T, U, V)
        loss = tsasr_b200.transducer_loss(logits, targets, il, tl, blank_index=0, reduction="mean", use_torchaudio=True)
        loss.backward()
        return loss.item(), [x.grad.clone() for x in (e, dc, jn.weight, jn.bias, head.weight, head.bias)]

    loss_f, grads_f = run("fused")
    loss_h, grads_h = run("by hand")
    loss_e, grads_e = run("eager")
    assert loss_f == loss_h and all(torch.equal(a, b_) for a, b_ in zip(grads_f, grads_h))
    assert abs(loss_f - loss_e) < 3e-4 * abs(loss_e)
    for a, b_ in zip(grads_f, grads_e):
        assert ((a - b_).norm() / b_.norm()).item() < 5e-2
