"""GPU parity tests of the projection GEMMs (encoder_proj / decoder_proj) through the C ABI.

Oracle: oracle/projection.py (float64 restatement of SB/nnet/linear.py:74 and its autograd backward).  The kernels run on
the bf16 tensor cores with an in-kernel (hi, lo) split of the fp32 operands; the bar is the split's own worst case, elementwise:
4.6e-5 * sum_k |x_k||w_k| (3 * 2^-16 per term: the dropped lo*lo product and the rounding of each lo part; reached only when
every term errs the same way -- a K = 3 contraction measured 0.3 of it), plus, at every shape, a max error below 3e-5 of the
largest output (measured 1.4e-5 at the recipe's K = 256; plain bf16 operands would give 4e-3)."""
import numpy as np
import pytest
import torch

import tsasr_b200
from tsasr_b200 import _lib, linear as tlin
from oracle import projection as oracle

pytestmark = pytest.mark.gpu
REL = 4.6e-5


def _dev():
    return torch.device("cuda:0")


def _call_fwd(x, w, b, want32=True, want16=True):
    d = _dev()
    R, K = x.shape
    N = w.shape[0]
    xd, wd = x.to(d), w.to(d)
    bd = b.to(d) if b is not None else None
    y = torch.empty((R, N), dtype=torch.float32, device=d) if want32 else None
    y16 = torch.empty((R, N), dtype=torch.bfloat16, device=d) if want16 else None
    _lib.check(_lib.load().tsasr_linear_fwd(xd.data_ptr(), wd.data_ptr(), bd.data_ptr() if bd is not None else None, R, K, N,
                                            y.data_ptr() if want32 else None, y16.data_ptr() if want16 else None,
                                            torch.cuda.current_stream(d).cuda_stream))
    torch.cuda.synchronize()
    return y, y16


SHAPES = [(6400, 256, 640), (1600, 512, 640), (137, 100, 29), (1, 7, 5), (300, 64, 128), (129, 65, 257), (5, 1000, 3), (260, 12, 132)]


@pytest.mark.parametrize("R,K,N", SHAPES)
@pytest.mark.parametrize("bias", [True, False])
def test_linear_fwd_matches_oracle(R, K, N, bias):
    g = torch.Generator().manual_seed(R + K + N)
    x = torch.randn(R, K, generator=g)
    w = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    b = (torch.rand(N, generator=g) * 2 - 1) if bias else None
    y, y16 = _call_fwd(x, w, b)
    ref = oracle.linear_fwd(x.numpy(), w.numpy(), None if b is None else b.numpy())
    bound = REL * oracle.linear_abs_bound(x.numpy(), w.numpy(), None if b is None else b.numpy()) + 1e-30
    err = np.abs(y.cpu().double().numpy() - ref)
    assert (err <= bound).all(), f"worst ratio {np.max(err / bound):.3f}"
    # beside the elementwise bound: max error relative to the largest output (the (hi, lo) split keeps ~17 bits per term)
    assert err.max() <= 3e-5 * np.abs(ref).max()
    # the bf16 output is exactly the rounding of the fp32 output: it is what the fused loss would compute itself
    assert torch.equal(y16, y.bfloat16())


def test_linear_fwd_single_outputs():
    g = torch.Generator().manual_seed(3)
    x, w, b = torch.randn(200, 96, generator=g), torch.randn(72, 96, generator=g) * 0.1, torch.randn(72, generator=g)
    y, y16 = _call_fwd(x, w, b)
    y_only, _ = _call_fwd(x, w, b, want16=False)
    _, y16_only = _call_fwd(x, w, b, want32=False)
    assert torch.equal(y, y_only) and torch.equal(y16, y16_only)


@pytest.mark.parametrize("R,K,N", SHAPES)
def test_linear_bwd_matches_oracle(R, K, N):
    g = torch.Generator().manual_seed(7 * R + K + N)
    d = _dev()
    x = torch.randn(R, K, generator=g)
    w = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    dy = torch.randn(R, N, generator=g) * 0.3
    lib = _lib.load()
    ws = torch.empty((lib.tsasr_linear_bwd_workspace_bytes(R, K, N) + 512,), dtype=torch.uint8, device=d)
    outs = {k: torch.full(s, float("nan"), device=d) for k, s in (("dx", (R, K)), ("dw", (N, K)), ("db", (N,)))}
    xd, wd, dyd = x.to(d), w.to(d), dy.to(d)
    _lib.check(lib.tsasr_linear_bwd(dyd.data_ptr(), xd.data_ptr(), wd.data_ptr(), R, K, N, outs["dx"].data_ptr(), outs["dw"].data_ptr(),
                                    outs["db"].data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(d).cuda_stream))
    torch.cuda.synchronize()
    rdx, rdw, rdb = oracle.linear_bwd(dy.numpy(), x.numpy(), w.numpy())
    a_dy, a_x, a_w = np.abs(dy.double().numpy()), np.abs(x.double().numpy()), np.abs(w.double().numpy())
    for name, ref, bound in (("dx", rdx, a_dy @ a_w), ("dw", rdw, a_dy.T @ a_x), ("db", rdb, a_dy.sum(0))):
        got = outs[name].cpu().double().numpy()
        assert np.isfinite(got).all(), name
        err = np.abs(got - ref)
        # long contractions (dW, db over R rows): fp32 accumulation of R terms
        lim = (REL + 6e-8 * np.sqrt(R)) * bound + 1e-30
        assert (err <= lim).all(), f"{name}: worst ratio {np.max(err / lim):.3f}"
    # dX / db can be skipped
    dw2 = torch.empty_like(outs["dw"])
    _lib.check(lib.tsasr_linear_bwd(dyd.data_ptr(), xd.data_ptr(), wd.data_ptr(), R, K, N, None, dw2.data_ptr(), None, ws.data_ptr(),
                                    ws.numel(), torch.cuda.current_stream(d).cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(dw2, outs["dw"])  # deterministic: same split-K fold order


def test_linear_module_is_a_dropin_for_nn_linear():
    """tsasr_b200.Linear against torch's own fp32 Linear on the GPU (TF32 off: the reference's setting), forward and
    backward through autograd, 3-D input as encoder_proj sees it."""
    torch.backends.cuda.matmul.allow_tf32 = False
    d = _dev()
    torch.manual_seed(0)
    ours = tsasr_b200.Linear(640, input_size=256).to(d)
    ref = torch.nn.Linear(256, 640).to(d)
    ref.load_state_dict(ours.w.state_dict())
    assert [n for n, _ in ours.named_parameters()] == ["w.weight", "w.bias"]  # checkpoint names of SB/nnet/linear.py:61
    x = torch.randn(4, 50, 256, device=d)
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    ya, yb = ours(xa), ref(xb)
    assert ya.shape == yb.shape == (4, 50, 640)
    gy = torch.randn_like(ya)
    ya.backward(gy)
    yb.backward(gy)
    for got, want in ((ya, yb), (xa.grad, xb.grad), (ours.w.weight.grad, ref.weight.grad), (ours.w.bias.grad, ref.bias.grad)):
        assert (got - want).abs().max().item() <= 2e-5 * max(want.abs().max().item(), 1.0)
    # no bias, frozen input
    ours_nb = tsasr_b200.Linear(32, input_shape=[None, None, 48], bias=False).to(d)
    y = ours_nb(torch.randn(3, 5, 48, device=d))
    y.sum().backward()
    assert ours_nb.w.weight.grad is not None and ours_nb.w.bias is None


def test_bf16_twin_feeds_the_fused_loss_bit_identically():
    """encoder_proj / decoder_proj through tsasr_b200.Linear: the joint consumes the bf16 copies the projection epilogue
    wrote (no cast pass), with results bit-identical to handing it the same fp32 tensors without their twins."""
    d = _dev()
    torch.manual_seed(1)
    B, T, U, H, V = 3, 40, 12, 640, 300
    enc_proj, dec_proj = tsasr_b200.Linear(H, input_size=256).to(d), tsasr_b200.Linear(H, input_size=512).to(d)
    head = torch.nn.Linear(H, V).to(d)
    joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
    xe, xd = torch.randn(B, T, 256, device=d), torch.randn(B, U, 512, device=d)
    targets = torch.randint(1, V, (B, U - 1), device=d)
    il, tl = torch.tensor([1.0, 0.8, 0.55], device=d), torch.tensor([1.0, 0.5, 0.7], device=d)

    def run(drop_twins):
        for m in (enc_proj, dec_proj, head):
            m.zero_grad()
        e, p = enc_proj(xe), dec_proj(xd)
        assert tlin.bf16_twin(e) is not None and tlin.bf16_twin(p) is not None
        assert torch.equal(tlin.bf16_twin(e), e.detach().bfloat16())
        if drop_twins:
            tlin.forget_twins()
            assert tlin.bf16_twin(e) is None
        n0 = _lib.launch_count()
        loss = tsasr_b200.transducer_loss(head(joiner(e[..., None, :], p[:, None, ...])), targets, il, tl, blank_index=0)
        loss.backward()
        torch.cuda.synchronize()
        return loss.detach().clone(), [m.weight.grad.clone() for m in (enc_proj.w, dec_proj.w, head)], _lib.launch_count() - n0

    l1, g1, _ = run(False)
    l2, g2, _ = run(True)
    assert torch.equal(l1, l2)
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)
    # an in-place update of the fp32 output invalidates its twin
    e = enc_proj(xe)
    e.mul_(2.0)
    assert tlin.bf16_twin(e) is None
    # a partial view is not the whole operand
    e = enc_proj(xe)
    assert tlin.bf16_twin(e[1:]) is None


def test_linear_fuzz_small_shapes_vs_oracle():
    """30 random shapes (ragged tiles, K tails, scalar and vector paths, split-K and single-pass dW): forward and backward
    against the float64 oracle."""
    rng = np.random.default_rng(2024)
    d = _dev()
    lib = _lib.load()
    st = torch.cuda.current_stream(d).cuda_stream
    for case in range(30):
        R, K, N = int(rng.integers(1, 400)), int(rng.integers(1, 260)), int(rng.integers(1, 300))
        if case % 5 == 0:
            R = int(rng.integers(1000, 5000))  # enough k-blocks for a split-K dW
        g = torch.Generator().manual_seed(case)
        x, w = torch.randn(R, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
        b = torch.randn(N, generator=g) if case % 2 else None
        dy = torch.randn(R, N, generator=g)
        y, y16 = _call_fwd(x, w, b)
        ref = oracle.linear_fwd(x.numpy(), w.numpy(), None if b is None else b.numpy())
        bound = REL * oracle.linear_abs_bound(x.numpy(), w.numpy(), None if b is None else b.numpy()) + 1e-30
        assert (np.abs(y.cpu().double().numpy() - ref) <= bound).all(), (case, R, K, N)
        assert torch.equal(y16, y.bfloat16()), (case, R, K, N)
        xd, wd, dyd = x.to(d), w.to(d), dy.to(d)
        ws = torch.empty((lib.tsasr_linear_bwd_workspace_bytes(R, K, N) + 256,), dtype=torch.uint8, device=d)
        dx, dw, db = torch.empty(R, K, device=d), torch.empty(N, K, device=d), torch.empty(N, device=d)
        _lib.check(lib.tsasr_linear_bwd(dyd.data_ptr(), xd.data_ptr(), wd.data_ptr(), R, K, N, dx.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                        ws.data_ptr(), ws.numel(), st))
        torch.cuda.synchronize()
        rdx, rdw, rdb = oracle.linear_bwd(dy.numpy(), x.numpy(), w.numpy())
        a_dy, a_x, a_w = np.abs(dy.double().numpy()), np.abs(x.double().numpy()), np.abs(w.double().numpy())
        for name, got, want, bnd in (("dx", dx, rdx, a_dy @ a_w), ("dw", dw, rdw, a_dy.T @ a_x), ("db", db, rdb, a_dy.sum(0))):
            lim = (REL + 6e-8 * np.sqrt(R)) * bnd + 1e-30
            assert (np.abs(got.cpu().double().numpy() - want) <= lim).all(), (case, name, R, K, N)


@pytest.mark.parametrize("name", ["linear_proj", "linear_proj_combine", "linear_proj_nobias"])
def test_linear_dropin_matches_reference_golden(golden, name):
    """tsasr_b200.Linear against outputs and gradients produced by the reference's own speechbrain.nnet.linear.Linear
    (tests/golden/linear_proj*.npz): 3-D input, combine_dims on a 4-D input, no bias."""
    g = golden(name)
    d = _dev()
    x = torch.from_numpy(g["x"]).to(d).requires_grad_()
    has_bias = "bias" in g
    if int(g["combine_dims"]):
        lin = tsasr_b200.Linear(g["weight"].shape[0], input_shape=[None, None, *g["x"].shape[2:]], bias=has_bias, combine_dims=True).to(d)
    else:
        lin = tsasr_b200.Linear(g["weight"].shape[0], input_size=g["weight"].shape[1], bias=has_bias).to(d)
    with torch.no_grad():
        lin.w.weight.copy_(torch.from_numpy(g["weight"]))
        if has_bias:
            lin.w.bias.copy_(torch.from_numpy(g["bias"]))
    n0 = _lib.launch_count()
    y = lin(x)
    assert _lib.launch_count() == n0 + 1  # the tcgen05 kernel ran
    (y * torch.from_numpy(g["d_out"]).to(d)).sum().backward()
    torch.cuda.synchronize()
    for got, key in ((y, "out"), (x.grad, "d_x"), (lin.w.weight.grad, "d_weight")) + (((lin.w.bias.grad, "d_bias"),) if has_bias else ()):
        want = g[key]
        assert tuple(got.shape) == want.shape, key
        assert np.abs(got.detach().cpu().numpy() - want).max() <= 5e-5 * max(np.abs(want).max(), 1.0), key
