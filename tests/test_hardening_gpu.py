"""Robustness tests of the C ABI on the GPU (round-2 additions): out-of-range lengths on the entry points that do not
validate them, torchaudio's clamp order, guard bands around every output buffer (compute-sanitizer is closed on this
pool -- see profiles/r2_sanitizer_status.txt -- so overruns are looked for with canaries), early shape errors."""
import numpy as np
import pytest
import torch

import tsasr_b200
from tsasr_b200 import _lib, ops
from tsasr_b200.functional import NumbaSemanticsTransducer

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def test_out_of_range_lengths_on_unvalidated_entry_points_are_clamped():
    """Transducer.apply / TransducerLoss on dense logits / rnnt_loss(check_lengths=False) perform no host-side length
    validation (neither does the reference).  T_b > T or a label count > U-1 must behave exactly like the clamped
    lengths in EVERY kernel (DP and gradient on the same rectangle), never index outside the utterance's slab."""
    g = torch.Generator().manual_seed(3)
    B, T, U, V = 3, 11, 6, 17
    logits = torch.randn(B, T, U, V, generator=g)
    targets = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32)
    bad_T = torch.tensor([T + 7, T, 0], dtype=torch.int32)
    bad_U = torch.tensor([U - 1, U + 40, -2], dtype=torch.int32)
    ok_T = torch.tensor([T, T, 1], dtype=torch.int32)
    ok_U = torch.tensor([U - 1, U - 1, 0], dtype=torch.int32)
    d = _dev()
    outs = []
    for Tl, Ul in ((bad_T, bad_U), (ok_T, ok_U)):
        lp = logits.to(d).log_softmax(-1).requires_grad_()
        loss = NumbaSemanticsTransducer.apply(lp, targets.to(d), Tl.to(d), Ul.to(d), 0, "none")
        loss.sum().backward()  # the VALUE divides by the T the caller passed (transducer_loss.py:104-106): gradients are compared
        x = logits.to(d).requires_grad_()
        costs = tsasr_b200.rnnt_loss(x, targets.to(d), Tl.to(d), Ul.to(d), blank=0, reduction="none", check_lengths=False)
        costs.sum().backward()
        torch.cuda.synchronize()
        outs.append((lp.grad.cpu(), costs.detach().cpu(), x.grad.cpu()))
    (g_bad, c_bad, x_bad), (g_ok, c_ok, x_ok) = outs
    assert torch.equal(g_bad, g_ok) and torch.equal(c_bad, c_ok) and torch.equal(x_bad, x_ok)
    assert torch.isfinite(c_ok).all()


@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_clamp_is_applied_to_the_unit_gradient_like_torchaudio(reduction):
    """torchaudio clamps the gradient of the UNIT cost and multiplies by grad_output afterwards
    (functional.py:1729-1734): with reduction="mean" the threshold must not become B times looser."""
    from torchaudio.functional import rnnt_loss as ta_rnnt_loss

    g = torch.Generator().manual_seed(9)
    B, T, U, V = 4, 13, 5, 11
    logits = 3.0 * torch.randn(B, T, U, V, generator=g)
    targets = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32)
    ll = torch.tensor([T, T - 2, T - 5, T], dtype=torch.int32)
    tl = torch.tensor([U - 1, 2, U - 1, 1], dtype=torch.int32)
    clamp = 0.05
    ref_x = logits.clone().requires_grad_()
    ref = ta_rnnt_loss(ref_x, targets, ll, tl, blank=0, clamp=clamp, reduction=reduction)
    (ref.sum() if ref.dim() else ref).backward()
    d = _dev()
    x = logits.to(d).requires_grad_()
    got = tsasr_b200.rnnt_loss(x, targets.to(d), ll.to(d), tl.to(d), blank=0, clamp=clamp, reduction=reduction)
    (got.sum() if got.dim() else got).backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-4)
    scale = 1.0 / B if reduction == "mean" else 1.0
    assert ref_x.grad.abs().max().item() == pytest.approx(clamp * scale, rel=1e-6)  # the clamp is active in this case
    assert (x.grad.cpu() - ref_x.grad).abs().max().item() < 1e-3 * scale + 1e-6


def _guarded(n_bytes, guard=4096):
    """(whole uint8 buffer, 1024-aligned payload view of n_bytes) with `guard` canary bytes on either side."""
    buf = torch.full((n_bytes + 2 * guard + 2048,), 0xA5, dtype=torch.uint8, device=_dev())
    start = guard + (-(buf.data_ptr() + guard)) % 1024
    return buf, buf[start: start + n_bytes], start


def _guards_intact(buf, start, n_bytes):
    return bool((buf[:start] == 0xA5).all()) and bool((buf[start + n_bytes:] == 0xA5).all())


@pytest.mark.parametrize("shape", [(2, 24, 9, 64, 40), (3, 37, 19, 128, 300), (2, 50, 33, 640, 1031), (1, 9, 130, 192, 257)])
@pytest.mark.parametrize("chunk", [0, 128 * 3])
def test_no_kernel_writes_outside_its_buffers(shape, chunk):
    """Every output and the workspace of the fused path sit between canary bands (raw C-ABI calls, exact sizes from
    the header's contracts): forward, DP and a chunked backward must leave all bands intact."""
    B, T, U, H, V = shape
    lib = _lib.load()
    d = _dev()
    g = torch.Generator().manual_seed(sum(shape))
    enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().to(d)
    dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().to(d)
    W = ((torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5).bfloat16().to(d)
    bias = ((torch.rand(V, generator=g) * 2 - 1) / H ** 0.5).to(d)
    targets = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).to(d)
    ll = torch.randint(max(1, T // 2), T + 1, (B,), generator=g, dtype=torch.int32)
    tl = torch.randint(0, U, (B,), generator=g, dtype=torch.int32)
    ll[0], tl[0] = T, U - 1
    ll, tl = ll.to(d), tl.to(d)
    dcost = torch.ones(B, device=d)
    n = ops.lattice_elems(B, T, U)
    st = torch.cuda.current_stream(d).cuda_stream
    ws_bytes = int(lib.tsasr_joint_bwd_workspace_bytes(B, T, U, H, V, chunk))
    sizes = {"lat2": 8 * n, "logz": 4 * n, "alpha": 4 * n, "beta": 4 * n, "cost": 4 * B, "lla": 4 * B, "llb": 4 * B,
             "d_enc": 4 * B * T * H, "d_dec": 4 * B * U * H, "dW": 4 * V * H, "db": 4 * V, "ws": ws_bytes}
    bufs = {k: _guarded(v) for k, v in sizes.items()}
    p = {k: v[1].data_ptr() for k, v in bufs.items()}
    _lib.check(lib.tsasr_joint_fwd(enc.data_ptr(), dec.data_ptr(), W.data_ptr(), bias.data_ptr(), targets.data_ptr(), ll.data_ptr(),
                                   tl.data_ptr(), B, T, U, H, V, 0, 0, 0.01, p["lat2"], p["logz"], st))
    _lib.check(lib.tsasr_lattice_alpha_beta(p["lat2"], ll.data_ptr(), tl.data_ptr(), B, T, U, p["alpha"], p["beta"], p["cost"],
                                            p["lla"], p["llb"], st))
    _lib.check(lib.tsasr_joint_bwd(enc.data_ptr(), dec.data_ptr(), W.data_ptr(), bias.data_ptr(), targets.data_ptr(), ll.data_ptr(),
                                   tl.data_ptr(), B, T, U, H, V, 0, 0, 0.01, p["lat2"], p["logz"], p["alpha"], p["beta"], p["cost"],
                                   dcost.data_ptr(), p["ws"], ws_bytes, chunk, -30.0, -1.0, p["d_enc"], p["d_dec"], p["dW"], p["db"], st))
    torch.cuda.synchronize()
    for k, (buf, view, start) in bufs.items():
        assert _guards_intact(buf, start, sizes[k]), f"canary band around {k} was overwritten"
    cost = bufs["cost"][1].view(torch.float32)
    assert torch.isfinite(cost).all() and (cost > 0).all()
    assert torch.isfinite(bufs["dW"][1].view(torch.float32)).all()


def test_unsupported_shapes_fail_where_the_call_is_made():
    d = _dev()
    U = 8200  # wider than the DP's shared-memory diagonal buffers allow (kMaxLatticeWidth = 8192)
    enc = torch.zeros(1, 4, 64, device=d)
    dec = torch.zeros(1, U, 64, device=d)
    W = torch.zeros(8, 64, device=d)
    with pytest.raises(NotImplementedError, match="lattice width"):
        tsasr_b200.fused_joint_rnnt_loss(enc, dec, W, None, torch.zeros(1, U - 1, dtype=torch.int32, device=d),
                                         torch.tensor([4], dtype=torch.int32, device=d), torch.tensor([U - 1], dtype=torch.int32, device=d))
    with pytest.raises(NotImplementedError, match="lattice width"):
        tsasr_b200.rnnt_loss(torch.zeros(1, 2, U, 3, device=d), torch.zeros(1, U - 1, dtype=torch.int32, device=d),
                             torch.tensor([2], dtype=torch.int32, device=d), torch.tensor([U - 1], dtype=torch.int32, device=d), check_lengths=False)
    # a head with a single output (blank only) is not deferred: the handle materialises with the eager math
    joiner = tsasr_b200.Transducer_joint()
    h = joiner(torch.zeros(1, 3, 1, 64, device=d), torch.zeros(1, 1, 2, 64, device=d))
    out = torch.nn.functional.linear(h, torch.zeros(1, 64, device=d))
    assert not isinstance(out, tsasr_b200.JointHandle) and out.shape == (1, 3, 2, 1)


@pytest.mark.parametrize("T,U", [(7, 1500), (40, 1025), (3, 3000)])
def test_wide_lattices_beyond_1024_columns_vs_torchaudio(T, U):
    """U > 1024 (more label positions than one thread block has threads): the DP switches to alpha_beta_wide_kernel
    (several columns per thread).  torchaudio's path has no such limit (the reference's Numba path does); loss and dlogits
    against torchaudio's CPU rnnt_loss, compat and fused paths."""
    from torchaudio.functional import rnnt_loss as ta_rnnt_loss

    g = torch.Generator().manual_seed(T * U)
    B, V = 2, 6
    logits = torch.randn(B, T, U, V, generator=g)
    targets = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32)
    ll = torch.tensor([T, max(1, T - 2)], dtype=torch.int32)
    tl = torch.tensor([U - 1, U - 300], dtype=torch.int32)
    ref_x = logits.clone().requires_grad_()
    ref = ta_rnnt_loss(ref_x, targets, ll, tl, blank=0, reduction="none")
    ref.sum().backward()
    d = _dev()
    x = logits.to(d).requires_grad_()
    got = tsasr_b200.rnnt_loss(x, targets.to(d), ll.to(d), tl.to(d), blank=0, reduction="none")
    got.sum().backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-4)
    assert (x.grad.cpu() - ref_x.grad).abs().max().item() < 1e-3
    # the two DP directions agree on log P through the C ABI
    lat2, den = ops.logits_to_lattice(logits.to(d), targets.to(d), ll.to(d), tl.to(d), 0)
    _, _, cost, ll_a, ll_b = ops.alpha_beta(lat2, ll.to(d), tl.to(d), B, T, U)
    np.testing.assert_allclose(ll_a.cpu().numpy(), ll_b.cpu().numpy(), rtol=1e-5)
    # fused path at the same width (H = 64, tiny head)
    gen = torch.Generator().manual_seed(1)
    enc = (0.5 * torch.randn(B, T, 64, generator=gen)).bfloat16()
    dec = (0.5 * torch.randn(B, U, 64, generator=gen)).bfloat16()
    W = ((torch.rand(V, 64, generator=gen) * 2 - 1) / 8).bfloat16()
    bias = torch.zeros(V)
    from oracle.reference_chain import reference_joint_loss_fwd_bwd

    costs = tsasr_b200.fused_joint_rnnt_loss(enc.to(d).float(), dec.to(d).float(), W.to(d).float(), bias.to(d), targets.to(d), ll.to(d), tl.to(d),
                                             blank=0, reduction="none")
    want = reference_joint_loss_fwd_bwd(enc, dec, W, bias, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=True)
    np.testing.assert_allclose(costs.cpu().numpy(), want["costs"].numpy(), rtol=1e-4)


def test_two_devices_in_one_process_use_their_own_launch_attributes():
    """cudaFuncSetAttribute is per device: a wide lattice (ring > 48 KB of dynamic shared memory) must
    launch on a second GPU of the same process."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    B, T, U, V = 1, 3, 900, 5
    for idx in (0, 1):
        d = torch.device("cuda", idx)
        logits = torch.zeros(B, T, U, V, device=d)
        costs = tsasr_b200.rnnt_loss(logits, torch.ones(B, U - 1, dtype=torch.int32, device=d),
                                     torch.tensor([T], dtype=torch.int32, device=d), torch.tensor([U - 1], dtype=torch.int32, device=d), blank=0,
                                     reduction="none")
        assert torch.isfinite(costs).all()


@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_fused_path_clamp_matches_the_compat_kernels(reduction):
    """torchaudio's ``clamp`` on the fused path (MODE_GRAD_CLAMP instantiation of the gradient pass): same operand
    gradients as the compat kernels with the same clamp on the materialised logits (those are pinned on torchaudio above),
    and the clamp does bite in this case."""
    from oracle.reference_chain import reference_joint_logits

    d = _dev()
    g = torch.Generator().manual_seed(31)
    B, T, U, H, V = 3, 37, 14, 128, 300
    enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().float()
    dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().float()
    W = (4.0 * (torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5).bfloat16().float()   # peaky softmax: large dlogits entries
    b = (torch.rand(V, generator=g) * 2 - 1) / H ** 0.5
    targets = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32)
    ll = torch.tensor([T, T - 9, T - 20], dtype=torch.int32)
    tl = torch.tensor([U - 1, 5, U - 1], dtype=torch.int32)
    clamp = 0.02
    outs = {}
    for c in (clamp, -1.0):
        e, dc, w, bb = (x.to(d).requires_grad_() for x in (enc, dec, W, b))
        loss = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, reduction=reduction, clamp=c)
        loss.backward()
        outs[c] = [x.grad.clone() for x in (e, dc, w, bb)]
    e2, dc2, w2, bb2 = (x.to(d).requires_grad_() for x in (enc, dec, W, b))
    logits = reference_joint_logits(e2, dc2, w2, bb2, "leaky_relu", 0.01, round_bf16=True)
    loss2 = tsasr_b200.rnnt_loss(logits.contiguous(), targets.to(d), ll.to(d), tl.to(d), blank=0, clamp=clamp, reduction=reduction)
    loss2.backward()
    for got, ref, free in zip(outs[clamp], (e2.grad, dc2.grad, w2.grad, bb2.grad), outs[-1.0]):
        assert ((got - ref).abs().max() / ref.abs().max()).item() < 3e-3
        assert ((got - free).abs().max() / free.abs().max()).item() > 5e-2   # the clamp changed the gradients


def test_deferred_finite_check_reports_one_step_late_without_syncing():
    """tsasr_b200.monitor: the check of step i is resolved when step i+1's check runs; a non-finite loss
    reaches the reference's own bookkeeping (here a stub of Brain.check_gradients, SB/core.py:1115-1150) one step later."""
    d = _dev()

    class Brain:
        def __init__(self):
            self.seen = []

        def check_gradients(self, loss):
            self.seen.append(float(loss))
            return bool(loss.isfinite())

    b = Brain()
    chk = tsasr_b200.monitor.install(b)
    assert b.check_gradients(torch.tensor(1.0, device=d)) is True            # nothing pending yet
    assert b.check_gradients(torch.tensor(float("nan"), device=d)) is True   # resolves step 0 (finite); NaN is pending
    assert b.seen == []                                                       # the original check has not run (no sync path)
    assert b.check_gradients(torch.tensor(2.0, device=d)) is False           # resolves the NaN step through the original check
    assert len(b.seen) == 1 and b.seen[0] != b.seen[0]
    assert chk.flush() is True and chk.deferred_steps == 3
    tsasr_b200.monitor.uninstall(b)
    assert b.check_gradients(torch.tensor(3.0, device=d)) is True and b.seen[-1] == 3.0


def _guarded_f32(n_elems):
    buf, view, start = _guarded(4 * n_elems)
    return buf, view.view(torch.float32), start, 4 * n_elems


@pytest.mark.parametrize("R,K,N", [(137, 100, 29), (129, 65, 257), (300, 64, 128), (1600, 512, 640)])
def test_canaries_around_the_projection_kernels(R, K, N):
    """tsasr_linear_fwd / tsasr_linear_bwd write exactly their outputs (ragged tiles, scalar and vector store paths,
    split-K workspace): guard bands around Y, Y_bf16, dX, dW, db and the workspace stay intact."""
    d = _dev()
    lib = _lib.load()
    g = torch.Generator().manual_seed(R + N)
    x, w, b = torch.randn(R, K, generator=g).to(d), (0.1 * torch.randn(N, K, generator=g)).to(d), torch.randn(N, generator=g).to(d)
    dy = torch.randn(R, N, generator=g).to(d)
    st = torch.cuda.current_stream(d).cuda_stream
    bufs = {}
    for name, n in (("y", R * N), ("dx", R * K), ("dw", N * K), ("db", N)):
        bufs[name] = _guarded_f32(n)
    y16_buf, y16_view, y16_start = _guarded(2 * R * N)
    ws_n = int(lib.tsasr_linear_bwd_workspace_bytes(R, K, N))
    ws_buf, ws_view, ws_start = _guarded(ws_n)
    _lib.check(lib.tsasr_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), R, K, N, bufs["y"][1].data_ptr(), y16_view.data_ptr(), st))
    _lib.check(lib.tsasr_linear_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), R, K, N, bufs["dx"][1].data_ptr(), bufs["dw"][1].data_ptr(),
                                    bufs["db"][1].data_ptr(), ws_view.data_ptr(), ws_n, st))
    torch.cuda.synchronize()
    for name, (buf, view, start, nbytes) in bufs.items():
        assert _guards_intact(buf, start, nbytes), name
        assert torch.isfinite(view).all(), name
    assert _guards_intact(y16_buf, y16_start, 2 * R * N) and _guards_intact(ws_buf, ws_start, ws_n)
    ref = torch.nn.functional.linear(x, w, b)
    assert (bufs["y"][1].view(R, N) - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()


@pytest.mark.parametrize("B,U,V,Hd", [(5, 7, 13, 128), (17, 9, 40, 256), (33, 5, 21, 512)])
def test_canaries_around_the_predictor_kernels(B, U, V, Hd):
    """tsasr_lstm_fwd / tsasr_lstm_bwd / tsasr_onehot_dw with batch sizes that leave the last batch tile partly empty:
    every output (incl. the sentinel-filled hand-off buffers) is written exactly, nothing else is touched."""
    d = _dev()
    lib = _lib.load()
    g = torch.Generator().manual_seed(B + U)
    G = 4 * Hd
    tokens = torch.randint(0, V, (B, U), generator=g, dtype=torch.int32).to(d)
    W_ih = (0.1 * torch.randn(G, V - 1, generator=g)).to(d)
    W_hh = (0.05 * torch.randn(G, Hd, generator=g)).to(d)
    b_ih, b_hh = (0.1 * torch.randn(G, generator=g)).to(d), (0.1 * torch.randn(G, generator=g)).to(d)
    rel = (torch.rand(B, generator=g) * 0.8 + 0.2).to(d)
    d_out = torch.randn(B, U, Hd, generator=g).to(d)
    st = torch.cuda.current_stream(d).cuda_stream
    bufs = {}
    for name, n in (("out", B * U * Hd), ("hprev", B * U * Hd), ("gates", B * U * G), ("cells", B * U * Hd), ("h_n", B * Hd), ("c_n", B * Hd),
                    ("dG", B * U * G), ("dW_ih", G * (V - 1))):
        bufs[name] = _guarded_f32(n)
    L_buf, L_view, L_start = _guarded(4 * B)
    p = {k: v[1].data_ptr() for k, v in bufs.items()}
    _lib.check(lib.tsasr_lstm_fwd(tokens.data_ptr(), 0, 0, V - 1, None, W_ih.data_ptr(), W_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(),
                                  rel.data_ptr(), None, B, U, Hd, p["out"], p["hprev"], p["gates"], p["cells"], p["h_n"], p["c_n"],
                                  L_view.data_ptr(), st))
    _lib.check(lib.tsasr_lstm_bwd(d_out.data_ptr(), None, None, W_hh.data_ptr(), p["gates"], p["cells"], L_view.data_ptr(), B, U, Hd, p["dG"], st))
    _lib.check(lib.tsasr_onehot_dw(tokens.data_ptr(), 0, 0, V - 1, p["dG"], B * U, G, p["dW_ih"], st))
    torch.cuda.synchronize()
    L = L_view.view(torch.int32).cpu()
    assert L.tolist() == (rel.cpu() * U).to(torch.int64).tolist()
    for name, (buf, view, start, nbytes) in bufs.items():
        assert _guards_intact(buf, start, nbytes), name
    assert _guards_intact(L_buf, L_start, 4 * B)
    # no sentinel (NaN) survives in the hand-off buffers; padded positions are exact zeros
    out, dG = bufs["out"][1].view(B, U, Hd), bufs["dG"][1].view(B, U, G)
    assert torch.isfinite(out).all() and torch.isfinite(dG).all() and torch.isfinite(bufs["dW_ih"][1]).all()
    for b_ in range(B):
        assert not out[b_, int(L[b_]):].any() and not dG[b_, int(L[b_]):].any()
    # positions whose gates were never written (padding) are never read: gates / cells there still hold the canary fill
    torch.backends.cudnn.allow_tf32 = False  # the reference recurrence in fp32, like the kernels (cuDNN's default is TF32)
    ref = torch.nn.LSTM(V - 1, Hd, batch_first=True).to(d)
    with torch.no_grad():
        ref.weight_ih_l0.copy_(W_ih); ref.weight_hh_l0.copy_(W_hh); ref.bias_ih_l0.copy_(b_ih); ref.bias_hh_l0.copy_(b_hh)
        onehot = torch.zeros(B, U, V - 1, device=d)
        nz = tokens != 0
        onehot[nz, (tokens[nz] - 1).long()] = 1.0
        ref_out = ref(onehot)[0]
    for b_ in range(B):
        assert (out[b_, :int(L[b_])] - ref_out[b_, :int(L[b_])]).abs().max().item() < 2e-5 if int(L[b_]) else True
