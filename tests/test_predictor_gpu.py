"""GPU parity tests of the prediction network kernels through the drop-in modules / the C ABI.

Oracle: oracle/predictor.py (float64 restatement of SB/nnet/embedding.py:76-114 + SB/nnet/RNN.py:25-54,254-278), pinned
against the reference's own modules by tests/golden/predictor_*.npz.  The kernels compute in fp32 like the reference
(cuDNN); the bar is agreement with the float64 oracle to fp32 round-off accumulated over the recurrence:
|got - ref| <= 2e-5 + 2e-5 |ref| for the states, 1e-4 relative to the largest entry for weight gradients (sums over B*U)."""
import numpy as np
import pytest
import torch

import tsasr_b200
from tsasr_b200 import _lib
from tsasr_b200.predictor import Embedding, LSTM, OneHotHandle
from oracle import predictor as oracle

pytestmark = pytest.mark.gpu
PREDICTOR_GOLDEN = ["predictor_onehot_ragged", "predictor_onehot_blank3", "predictor_onehot_full", "predictor_dense"]


def _dev():
    return torch.device("cuda:0")


def _close(got, ref, atol, rtol, what):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    lim = atol + rtol * np.abs(ref)
    assert (err <= lim).all(), f"{what}: worst ratio {np.max(err / lim):.3f}, max abs err {err.max():.3e}"


def _grad_close(got, ref, what, rel=1e-4):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    assert np.abs(got - ref).max() <= rel * scale, f"{what}: max err / max = {np.abs(got - ref).max() / scale:.3e}"


def _load_lstm(lstm, params):
    with torch.no_grad():
        lstm.rnn.weight_ih_l0.copy_(params["weight_ih"])
        lstm.rnn.weight_hh_l0.copy_(params["weight_hh"])
        lstm.rnn.bias_ih_l0.copy_(params["bias_ih"])
        lstm.rnn.bias_hh_l0.copy_(params["bias_hh"])


def _run_dropins(tokens, V, blank, params, rel, d_out, x_dense=None):
    d = _dev()
    Hd = params["weight_hh"].shape[1]
    if x_dense is None:
        emb = Embedding(num_embeddings=V, consider_as_one_hot=True, blank_id=blank).to(d)
        lstm = LSTM(input_shape=[None, None, V - 1], hidden_size=Hd, num_layers=1).to(d)
        x = emb(tokens.to(d))
        assert isinstance(x, OneHotHandle) and tuple(x.shape) == (*tokens.shape, V - 1)
    else:
        lstm = LSTM(input_size=x_dense.shape[-1], hidden_size=Hd, num_layers=1).to(d)
        x = x_dense.to(d).requires_grad_()
    _load_lstm(lstm, params)
    n0 = _lib.launch_count()
    out, (h_n, c_n) = lstm(x, lengths=rel.to(d) if rel is not None else None)
    fwd_launches = _lib.launch_count() - n0
    (out * d_out.to(d)).sum().backward()
    torch.cuda.synchronize()
    r = lstm.rnn
    res = {"out": out.detach().cpu(), "h_n": h_n.detach().cpu()[0], "c_n": c_n.detach().cpu()[0], "d_weight_ih": r.weight_ih_l0.grad.cpu(),
           "d_weight_hh": r.weight_hh_l0.grad.cpu(), "d_bias_ih": r.bias_ih_l0.grad.cpu(), "d_bias_hh": r.bias_hh_l0.grad.cpu(),
           "fwd_launches": fwd_launches}
    if x_dense is not None:
        res["d_x"] = x.grad.cpu()
    return res


@pytest.mark.parametrize("name", PREDICTOR_GOLDEN)
def test_predictor_matches_reference_golden(golden, name):
    """The drop-in Embedding + LSTM against outputs and gradients produced by the reference's own modules."""
    g = golden(name)
    params = {k: torch.from_numpy(g[k]) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")}
    rel = torch.from_numpy(g["rel_lengths"]) if g["rel_lengths"].size else None
    dense = torch.from_numpy(g["x"]) if "x" in g else None
    res = _run_dropins(torch.from_numpy(g["tokens"]), int(g["vocab"]), int(g["blank"]), params, rel, torch.from_numpy(g["d_out"]), dense)
    assert res["fwd_launches"] == (2 if dense is not None else 1)  # the whole recurrence is one launch (+ the input GEMM)
    for key in ("out", "h_n", "c_n"):
        _close(res[key], g[key], 2e-5, 2e-5, key)
    for key in ("d_weight_ih", "d_weight_hh", "d_bias_ih", "d_bias_hh") + (("d_x",) if dense is not None else ()):
        _grad_close(res[key], g[key], key, rel=2e-4)  # the golden gradients are the reference's fp32 results
    # padded positions: exact zeros, like pad_packed_sequence
    if rel is not None:
        L = oracle.packed_lengths(rel, g["tokens"].shape[1])
        for b, l in enumerate(L.tolist()):
            assert not res["out"][b, l:].any()


@pytest.mark.parametrize("B,U,V,Hd,blank", [(16, 100, 1000, 512, 0), (40, 23, 60, 256, 5), (3, 2, 9, 128, 0), (64, 7, 33, 128, 32)])
def test_predictor_matches_oracle_at_recipe_shapes(B, U, V, Hd, blank):
    g = torch.Generator().manual_seed(B + U + V)
    tokens = torch.randint(0, V, (B, U), generator=g)
    tokens[:, 0] = blank
    k = 1.0 / Hd ** 0.5
    params = {"weight_ih": (torch.rand(4 * Hd, V - 1, generator=g) * 2 - 1) * k, "weight_hh": torch.nn.init.orthogonal_(torch.empty(4 * Hd, Hd)),
              "bias_ih": (torch.rand(4 * Hd, generator=g) * 2 - 1) * k, "bias_hh": (torch.rand(4 * Hd, generator=g) * 2 - 1) * k}
    rel = torch.rand(B, generator=g) * 0.7 + 0.3
    rel[0] = 1.0
    d_out = torch.randn(B, U, Hd, generator=g)
    res = _run_dropins(tokens, V, blank, params, rel, d_out)
    ref = oracle.predictor_fwd_bwd(tokens, V, blank, params, rel, d_out)
    for key in ("out", "h_n", "c_n"):
        _close(res[key], ref[key], 2e-5, 2e-5, key)
    for key in ("d_weight_ih", "d_weight_hh", "d_bias_ih", "d_bias_hh"):
        _grad_close(res[key], ref[key], key)
    assert torch.equal(res["d_bias_ih"], res["d_bias_hh"])


def test_predictor_lengths_truncate_like_the_reference_packing():
    """Bit-exact integer step: (rel * U) in fp32, truncated (SB/nnet/RNN.py:35 + pack_padded_sequence) -- NOT rounded."""
    d = _dev()
    U, Hd, V = 13, 128, 11
    rel = torch.tensor([1.0, 0.6, 0.55, 0.0769231, 0.9999999, 0.5384616], dtype=torch.float32)
    B = rel.shape[0]
    lib = _lib.load()
    tokens = torch.zeros((B, U), dtype=torch.int32, device=d)
    W_ih, W_hh = torch.zeros(4 * Hd, V - 1, device=d), torch.zeros(4 * Hd, Hd, device=d)
    out, L = torch.empty(B, U, Hd, device=d), torch.empty(B, dtype=torch.int32, device=d)
    reld = rel.to(d)
    _lib.check(lib.tsasr_lstm_fwd(tokens.data_ptr(), 0, 0, V - 1, None, W_ih.data_ptr(), W_hh.data_ptr(), None, None, reld.data_ptr(), None,
                                  B, U, Hd, out.data_ptr(), None, None, None, None, None, L.data_ptr(),
                                  torch.cuda.current_stream(d).cuda_stream))
    torch.cuda.synchronize()
    assert L.cpu().tolist() == oracle.packed_lengths(rel, U).tolist() == (rel * U).to(torch.int64).tolist()


def test_predictor_agrees_with_cudnn_and_falls_back_where_it_must():
    """Cross-check against torch.nn.LSTM (cuDNN) on the materialised one-hot input, and the cases that take the
    reference's path: an initial state (decode-time single steps) and CPU tensors."""
    d = _dev()
    torch.manual_seed(3)
    torch.backends.cudnn.allow_tf32 = False  # cuDNN's recurrence in fp32 (its default is TF32)
    B, U, V, Hd = 8, 20, 50, 512
    emb = Embedding(num_embeddings=V, consider_as_one_hot=True, blank_id=0).to(d)
    lstm = LSTM(input_shape=[None, None, V - 1], hidden_size=Hd).to(d)
    assert lstm.__class__.__name__ == "LSTM"  # SB/decoders/transducer.py:491-499 recognises recurrent layers by name
    assert [n for n, _ in lstm.named_parameters()] == ["rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0"]
    assert list(emb.state_dict()) == ["Embedding.weight"] and not emb.Embedding.weight.requires_grad
    tokens = torch.randint(0, V, (B, U), device=d)
    with torch.no_grad():
        n0 = _lib.launch_count()
        ours, (h_n, c_n) = lstm(emb(tokens))
        assert _lib.launch_count() - n0 == 1
        ref, (rh, rc) = lstm.rnn(emb(tokens).materialize())
        assert (ours - ref).abs().max().item() < 2e-5 and (h_n - rh).abs().max().item() < 2e-5 and (c_n - rc).abs().max().item() < 2e-5
        # decode-time step: hidden state given -> cuDNN path, no launch of ours
        n0 = _lib.launch_count()
        step, hid = lstm(emb(tokens[:, :1]), (rh, rc))
        assert _lib.launch_count() == n0 and step.shape == (B, 1, Hd)
    cpu_emb, cpu_lstm = Embedding(num_embeddings=9, consider_as_one_hot=True), LSTM(input_size=8, hidden_size=16)
    x = cpu_emb(torch.tensor([[0, 3, 8]]))
    assert not isinstance(x, OneHotHandle) and x.shape == (1, 3, 8) and x[0, 0].abs().sum() == 0 and x[0, 1, 2] == 1 and x[0, 2, 7] == 1
    assert cpu_lstm(x)[0].shape == (1, 3, 16)


def test_predictor_chain_into_the_fused_loss():
    """embedding -> decoder -> decoder_proj -> joiner -> head -> loss with every drop-in against the same chain built from
    torch's own modules (train_librispeechmix_scratch.py:122-135,158): loss and the predictor's weight gradients."""
    d = _dev()
    torch.manual_seed(5)
    B, T, U, V, H, Hd = 4, 30, 9, 40, 640, 128
    emb = Embedding(num_embeddings=V, consider_as_one_hot=True, blank_id=0).to(d)
    lstm = LSTM(input_shape=[None, None, V - 1], hidden_size=Hd).to(d)
    dec_proj = tsasr_b200.Linear(H, input_size=Hd).to(d)
    head = torch.nn.Linear(H, V).to(d)
    joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
    enc_out = (0.5 * torch.randn(B, T, H, device=d)).bfloat16().float()
    tokens = torch.randint(1, V, (B, U - 1), device=d)
    tokens_bos = torch.cat([torch.zeros(B, 1, dtype=torch.long, device=d), tokens], dim=1)
    il = torch.tensor([1.0, 0.9, 0.6, 0.8], device=d)
    tl = torch.tensor([1.0, 0.5, 0.75, 0.25], device=d)
    bos_l = torch.tensor([1.0, 5 / 9, 7 / 9, 3 / 9], device=d)  # (labels + 1) / U

    dec_out, _ = lstm(emb(tokens_bos), lengths=bos_l)
    logits = head(joiner(enc_out[..., None, :], dec_proj(dec_out)[:, None, ...]))
    loss = tsasr_b200.transducer_loss(logits, tokens, il, tl, blank_index=0)
    loss.backward()
    got = {n: p.grad.clone() for n, p in lstm.named_parameters()}

    # the same chain with stock torch modules (cuDNN LSTM on the packed one-hot tensor, eager joint, torchaudio loss on CPU)
    from oracle.reference_chain import reference_transducer_loss
    ref_lstm = torch.nn.LSTM(V - 1, Hd, batch_first=True)
    ref_lstm.load_state_dict({k[4:]: v.detach().cpu() for k, v in lstm.state_dict().items()})
    x = torch.nn.functional.embedding(tokens_bos.cpu(), emb.Embedding.weight.detach().cpu())
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, (bos_l.cpu() * U), batch_first=True, enforce_sorted=False)
    o, _ = torch.nn.utils.rnn.pad_packed_sequence(ref_lstm(packed)[0], batch_first=True)
    p = torch.nn.functional.linear(o, dec_proj.w.weight.detach().cpu(), dec_proj.w.bias.detach().cpu())
    p = p + (p.detach().bfloat16().float() - p.detach())  # the one rounding the fused joint applies to its operands
    joint = torch.nn.functional.leaky_relu(enc_out.cpu()[..., None, :] + p[:, None, ...], 0.01)
    joint = joint + (joint.detach().bfloat16().float() - joint.detach())
    W = head.weight.detach().cpu().bfloat16().float()
    ref_loss = reference_transducer_loss(torch.nn.functional.linear(joint, W, head.bias.detach().cpu()), tokens.cpu(), il.cpu(), tl.cpu(), 0)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item())
    for n, pr in ref_lstm.named_parameters():
        ref_g = pr.grad
        err = (got["rnn." + n].cpu() - ref_g).abs().max().item() / ref_g.abs().max().item()
        assert err < 1e-2, (n, err)  # bf16 dlogits noise of the fused loss, not of the predictor kernels


def test_predictor_fuzz_vs_oracle():
    """16 random configurations (batch sizes that leave batch tiles partly empty, up to four passes, every hidden size, tiny
    vocabularies, blank anywhere, utterances of one step, int32 and int64 tokens): forward and backward vs the float64 oracle."""
    rng = np.random.default_rng(7)
    for case in range(16):
        B, U = int(rng.integers(1, 65)), int(rng.integers(2, 24))
        V, Hd = int(rng.integers(2, 60)), int(rng.choice([128, 256, 512]))
        blank = int(rng.integers(0, V))
        g = torch.Generator().manual_seed(100 + case)
        tokens = torch.randint(0, V, (B, U), generator=g, dtype=torch.int32 if case % 2 else torch.int64)
        k = 1.0 / Hd ** 0.5
        params = {"weight_ih": (torch.rand(4 * Hd, V - 1, generator=g) * 2 - 1) * k, "weight_hh": (torch.rand(4 * Hd, Hd, generator=g) * 2 - 1) * k,
                  "bias_ih": (torch.rand(4 * Hd, generator=g) * 2 - 1) * k, "bias_hh": (torch.rand(4 * Hd, generator=g) * 2 - 1) * k}
        rel = torch.rand(B, generator=g) * (1.0 - 1.0 / U) + 1.0 / U + 1e-4   # at least one step each
        rel.clamp_(max=1.0)
        rel[int(rng.integers(0, B))] = 1.0
        d_out = torch.randn(B, U, Hd, generator=g)
        res = _run_dropins(tokens, V, blank, params, rel, d_out)
        ref = oracle.predictor_fwd_bwd(tokens, V, blank, params, rel, d_out)
        for key in ("out", "h_n", "c_n"):
            _close(res[key], ref[key], 2e-5, 2e-5, f"case {case} (B={B},U={U},V={V},Hd={Hd},blank={blank}) {key}")
        for key in ("d_weight_ih", "d_weight_hh", "d_bias_ih", "d_bias_hh"):
            _grad_close(res[key], ref[key], f"case {case} (B={B},U={U},V={V},Hd={Hd},blank={blank}) {key}")
