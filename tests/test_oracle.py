"""Pins the CPU oracle (numpy float64 + C restatement) to the reference.

Anchors: (1) the reference's known-answer test, vendor/speechbrain/tests/unittests/test_losses.py:109-152
(2.2478, Numba semantics; x T = 4.4957 torchaudio semantics); (2) tests/golden/*.npz produced by
the reference's own transducer_loss (both branches) in the authoring container
(oracle/make_golden.py); (3) torchaudio's CPU rnnt_loss run live.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import rnnt_c, rnnt_numpy as rn
from oracle.reference_chain import reference_rnnt_abs, reference_transducer_loss

from conftest import GOLDEN

TA_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*_torchaudio*.npz")))
NB_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*_numba*.npz")))


def test_known_answer_value():
    # test_losses.py:152: assert out_cost.item() == pytest.approx(2.2478, 0.0001)
    x = np.array([[[[.1, .6, .1, .1, .1], [.1, .1, .6, .1, .1], [.1, .1, .2, .8, .1]],
                   [[.1, .6, .1, .1, .1], [.1, .1, .2, .1, .1], [.7, .1, .2, .1, .1]]]], dtype=np.float32)
    lp = rn.log_softmax(rn.log_softmax(x))  # the test log-softmaxes, then transducer_loss does again
    loss, _ = rn.rnnt_numba(lp, np.array([[1, 2]]), [2], [2], 0, "mean")
    assert float(loss) == pytest.approx(2.2478, rel=1e-4)
    per_utt, _ = rnnt_c.rnnt_numba(lp, np.array([[1, 2]]), [2], [2], 0)
    assert float(per_utt[0]) == pytest.approx(2.2478, rel=1e-4)
    costs, _ = rn.rnnt_torchaudio(lp, np.array([[1, 2]]), [2], [2], 0)
    assert float(costs[0]) == pytest.approx(2.2478 * 2, rel=1e-4)  # loss_numba * T = loss_torchaudio


def _scale(reduction, B):
    return 1.0 / B if reduction == "mean" else 1.0


@pytest.mark.parametrize("name", TA_CASES)
def test_torchaudio_semantics_vs_golden(golden, name):
    g = golden(name)
    B, T, U, V = g["logits"].shape
    # integer length conversion, bit exact (losses.py:58-59)
    assert np.array_equal(rn.lengths_from_relative(g["input_rel"], T), g["input_abs"])
    assert np.array_equal(rn.lengths_from_relative(g["target_rel"], U - 1), g["target_abs"])
    red = str(g["reduction"])
    for impl in ("numpy", "c32", "c64"):
        if impl == "numpy":
            costs, grads = rn.rnnt_torchaudio(g["logits"], g["targets"], g["input_abs"], g["target_abs"], int(g["blank"]))
        else:
            costs, grads = rnnt_c.rnnt_torchaudio(g["logits"], g["targets"], g["input_abs"], g["target_abs"],
                                                  int(g["blank"]), fp32=(impl == "c32"))
        loss = rn.reduce_costs(costs.astype(np.float64), red)
        np.testing.assert_allclose(loss, g["loss"], rtol=1e-4, err_msg=impl)
        # the fp32 reference carries alpha/beta of magnitude ~|cost|; its own rounding error in
        # exp(alpha+beta-L) grows with |cost| * 2^-24, so the tolerance scales with it
        atol = 2e-5 * max(1.0, float(np.abs(costs).max()) / 50.0)
        np.testing.assert_allclose(grads * _scale(red, B), g["dlogits"], atol=atol, rtol=0, err_msg=impl)
        # structural properties: rows sum to zero over V; exact zeros outside the T_b x U_b rectangle
        assert np.abs(grads.sum(-1)).max() < 1e-4
        for b in range(B):
            assert not grads[b, g["input_abs"][b]:].any() and not grads[b, :, g["target_abs"][b] + 1:].any()


@pytest.mark.parametrize("name", NB_CASES)
def test_numba_semantics_vs_golden(golden, name):
    g = golden(name)
    lp = rn.log_softmax(g["logits"])
    red = str(g["reduction"])
    loss, glp = rn.rnnt_numba(lp, g["targets"], g["input_abs"], g["target_abs"], int(g["blank"]), red)
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-4)
    # the golden holds d loss / d logits = log_softmax backward of the stored log-prob gradient,
    # which is NOT scaled by 1/B for "mean" (transducer_loss.py:280-293)
    dlogits = glp - np.exp(lp) * glp.sum(-1, keepdims=True)
    np.testing.assert_allclose(dlogits, g["dlogits"], atol=2e-5, rtol=0)
    per_utt, gc = rnnt_c.rnnt_numba(lp, g["targets"], g["input_abs"], g["target_abs"], int(g["blank"]))
    np.testing.assert_allclose(gc, glp, atol=2e-5, rtol=0)
    np.testing.assert_allclose({"mean": per_utt.mean(), "sum": per_utt.sum(), "none": per_utt}[red], g["loss"], rtol=1e-4)


def test_numba_times_T_is_torchaudio(golden):
    nb, ta = golden("ragged_numba_none"), golden("ragged_torchaudio_none")
    np.testing.assert_allclose(nb["loss"] * nb["input_abs"], ta["loss"], rtol=1e-5)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_vs_live_torchaudio(seed):
    rng = np.random.default_rng(seed)
    B, T, U, V = 3, 15 + seed, 7, 21
    logits = (rng.standard_normal((B, T, U, V)) * 2).astype(np.float32)
    targets = rng.integers(0, V - 1, (B, U - 1)).astype(np.int32)
    ll = np.array([T, max(1, T - 4), 1], dtype=np.int32)
    tl = np.array([U - 1, 0, U - 3], dtype=np.int32)
    blank = V - 1
    t = torch.tensor(logits, requires_grad=True)
    c = reference_rnnt_abs(t, torch.tensor(targets), torch.tensor(ll), torch.tensor(tl), blank=blank)
    c.sum().backward()
    for costs, grads in (rn.rnnt_torchaudio(logits, targets, ll, tl, blank),
                         rnnt_c.rnnt_torchaudio(logits, targets, ll, tl, blank)):
        np.testing.assert_allclose(costs, c.detach().numpy(), rtol=1e-5)
        np.testing.assert_allclose(grads, t.grad.numpy(), atol=2e-5, rtol=0)


def test_reference_chain_relative_lengths(golden):
    g = golden("half_rounding_torchaudio")
    t = torch.tensor(g["logits"], requires_grad=True)
    loss = reference_transducer_loss(t, torch.tensor(g["targets"]).long(), torch.tensor(g["input_rel"]),
                                     torch.tensor(g["target_rel"]), int(g["blank"]), reduction="none")
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-6)


def test_joint_chain_vs_golden(golden):
    for name in ("joint_leaky", "joint_tanh", "joint_relu"):
        g = golden(name)
        act, red = str(g["act"]), str(g["reduction"])
        J, logits = rn.joint_logits(g["enc"], g["dec"], g["W"], g["b"], act, round_bf16=False)
        costs, dlogits = rn.rnnt_torchaudio(logits, g["targets"], g["input_abs"], g["target_abs"], 0)
        np.testing.assert_allclose(rn.reduce_costs(costs, red), g["loss"], rtol=1e-5)
        d_enc, d_dec, dW, db = rn.joint_backward(J, dlogits * _scale(red, len(costs)), g["W"], act, round_bf16=False)
        for got, key in ((d_enc, "d_enc"), (d_dec, "d_dec"), (dW, "dW"), (db, "db")):
            np.testing.assert_allclose(got, g[key], atol=5e-5, rtol=1e-4, err_msg=f"{name}:{key}")


def test_config1_c_oracle_vs_torchaudio():
    """BASELINE config 1 (B=4,T=200,U=40,V=1000) at reduced T so the CPU suite stays fast."""
    gen = torch.Generator().manual_seed(0)
    B, T, U, V = 4, 50, 40, 1000
    logits = torch.randn(B, T, U, V, generator=gen)
    targets = torch.randint(1, V, (B, U - 1), generator=gen, dtype=torch.int32)
    ll = torch.tensor([50, 45, 38, 31], dtype=torch.int32)
    tl = torch.tensor([39, 30, 21, 10], dtype=torch.int32)
    t = logits.clone().requires_grad_()
    c = reference_rnnt_abs(t, targets, ll, tl)
    c.sum().backward()
    costs, grads = rnnt_c.rnnt_torchaudio(logits.numpy(), targets.numpy(), ll.numpy(), tl.numpy(), 0)
    np.testing.assert_allclose(costs, c.detach().numpy(), rtol=1e-5)
    assert np.abs(grads - t.grad.numpy()).max() < 1e-4


# ---- prediction network: the restatement against the reference's own modules ----
PREDICTOR_GOLDEN = ["predictor_onehot_ragged", "predictor_onehot_blank3", "predictor_onehot_full", "predictor_dense"]


@pytest.mark.parametrize("name", PREDICTOR_GOLDEN)
def test_predictor_oracle_matches_reference_modules(golden, name):
    """oracle/predictor.py (explicit float64 loops) against the outputs and gradients of speechbrain's Embedding + LSTM
    (oracle/make_golden_predictor.py ran the REAL modules of /root/reference): pins the one-hot column mapping incl.
    blank != 0, the gate order, the packed-sequence semantics and the TRUNCATING length conversion of the packing."""
    from oracle import predictor

    g = golden(name)
    params = {k: torch.from_numpy(g[k]) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")}
    rel = torch.from_numpy(g["rel_lengths"]) if g["rel_lengths"].size else None
    dense = torch.from_numpy(g["x"]) if "x" in g else None
    res = predictor.predictor_fwd_bwd(torch.from_numpy(g["tokens"]), int(g["vocab"]), int(g["blank"]), params, rel,
                                      torch.from_numpy(g["d_out"]), x_dense=dense)
    if name == "predictor_onehot_ragged":
        assert res["lengths"].tolist() == [13, 7, 3, 7, 1]  # 0.6 * 13 = 7.8 is truncated (the loss would round it to 8)
    for key in ("out", "h_n", "c_n", "d_weight_ih", "d_weight_hh", "d_bias_ih", "d_bias_hh") + (("d_x",) if dense is not None else ()):
        np.testing.assert_allclose(res[key].numpy(), g[key], rtol=2e-4, atol=2e-6, err_msg=key)


# ---- projections: the restatement against the reference's own Linear class ----
PROJECTION_GOLDEN = ["linear_proj", "linear_proj_combine", "linear_proj_nobias"]


@pytest.mark.parametrize("name", PROJECTION_GOLDEN)
def test_projection_oracle_matches_reference_module(golden, name):
    """oracle/projection.py against outputs and gradients of speechbrain.nnet.linear.Linear (oracle/make_golden_projection.py
    ran the REAL class): last-dimension product, bias, and the combine_dims flattening of 4-D inputs."""
    from oracle import projection

    g = golden(name)
    x = g["x"]
    if int(g["combine_dims"]) and x.ndim == 4:
        x = x.reshape(x.shape[0], x.shape[1], -1)      # SB/nnet/linear.py:71-72
    bias = g["bias"] if "bias" in g else None
    np.testing.assert_allclose(projection.linear_fwd(x, g["weight"], bias), g["out"], rtol=1e-5, atol=1e-6)
    K = x.shape[-1]
    dx, dw, db = projection.linear_bwd(g["d_out"].reshape(-1, g["d_out"].shape[-1]), x.reshape(-1, K), g["weight"])
    np.testing.assert_allclose(dx.reshape(g["d_x"].shape), g["d_x"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dw, g["d_weight"], rtol=1e-5, atol=1e-5)
    if bias is not None:
        np.testing.assert_allclose(db, g["d_bias"], rtol=1e-5, atol=1e-5)
