"""Decode-time joint step: the fused kernel behind
``TransducerBeamSearcher._joint_forward_step`` (SB/decoders/transducer.py:375-384).

Oracle: the reference's eager chain (oracle/greedy_decode.py restates the searcher's greedy loop); golden vectors
come from the REAL searcher class run in the authoring container (oracle/make_golden_decode.py).
Tolerance: fp32 arithmetic on both sides, log-probs within 2e-5 absolute; decoded label sequences bit-exact."""
import types

import numpy as np
import pytest
import torch

import tsasr_b200
from tsasr_b200 import decode
from oracle.greedy_decode import ToyPredictor, eager_joint_step, greedy_decode


def _golden_modules(g, device):
    B, T, V, E, HID, H = (int(x) for x in g["dims"])
    pred = ToyPredictor(V, E, HID, H)
    pred.load_state_dict({k[5:]: torch.tensor(g[k]) for k in g if k.startswith("pred.")})
    head = torch.nn.Linear(H, V)
    with torch.no_grad():
        head.weight.copy_(torch.tensor(g["W"]))
        head.bias.copy_(torch.tensor(g["b"]))
    tjoint = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
    hyps, o = [], 0
    for n in g["hyp_lens"]:
        hyps.append([int(x) for x in g["hyp_flat"][o:o + int(n)]])
        o += int(n)
    return pred.to(device).eval(), head.to(device).eval(), tjoint, torch.tensor(g["tn"]).to(device), hyps


def test_oracle_greedy_loop_reproduces_reference_searcher(golden):
    """CPU: the restated loop with the eager step gives exactly the hypotheses of the reference's searcher class."""
    g = golden("greedy_decode")
    pred, head, tjoint, tn, hyps = _golden_modules(g, "cpu")
    got, scores = greedy_decode(tn, pred.layers(), eager_joint_step(tjoint, [head], torch.nn.LogSoftmax(dim=-1)))
    assert got == hyps
    np.testing.assert_allclose(np.exp(np.array(scores)).mean(), float(g["mean_exp_score"]), rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,V,act", [(1, 640, 1000, "leaky_relu"), (16, 640, 1000, "leaky_relu"), (5, 64, 29, "tanh"),
                                       (40, 256, 5003, "relu"), (3, 128, 7, "identity")])
def test_decode_step_vs_eager_chain(B, H, V, act):
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 1000 + V)
    enc = torch.randn(B, H, generator=g).to(d)
    dec = torch.randn(B, H, generator=g).to(d)
    W = (torch.randn(V, H, generator=g) / H ** 0.5 * 3).to(d)
    b = torch.randn(V, generator=g).to(d)
    got = decode.joint_decode_step(enc, dec, W, b, act, 0.01)
    f = {"leaky_relu": lambda x: torch.nn.functional.leaky_relu(x, 0.01), "relu": torch.relu, "tanh": torch.tanh,
         "identity": lambda x: x}[act]
    ref = torch.log_softmax(torch.nn.functional.linear(f(enc.double() + dec.double()), W.double(), b.double()), dim=-1)
    assert (got.double() - ref).abs().max().item() < 2e-5
    top2 = ref.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4           # arg-max must agree wherever it is not a numerical tie
    assert torch.equal(got.argmax(-1)[clear], ref.argmax(-1)[clear])
    np.testing.assert_allclose(got.exp().sum(-1).cpu().numpy(), 1.0, rtol=1e-5)


@pytest.mark.gpu
def test_greedy_decode_with_fused_step_vs_reference_golden(golden):
    g = golden("greedy_decode")
    d = torch.device("cuda:0")
    pred, head, tjoint, tn, hyps = _golden_modules(g, d)
    softmax = torch.nn.LogSoftmax(dim=-1)
    step = decode.fused_joint_forward_step(tjoint, [head], softmax)
    assert step is not None
    torch.backends.cudnn.allow_tf32 = False  # the toy LSTM must run in fp32 like the CPU run that made the vector
    launches0 = tsasr_b200._lib.launch_count()
    got, scores = greedy_decode(tn, pred.layers(), step)
    assert tsasr_b200._lib.launch_count() - launches0 == 2 * tn.shape[1]   # logits + normalise kernel per decoded frame
    assert got == hyps
    np.testing.assert_allclose(np.exp(np.array(scores)).mean(), float(g["mean_exp_score"]), rtol=1e-4)
    # the first frames' full log-prob rows, as produced by the reference's own _joint_forward_step
    with torch.no_grad():
        cpu_pred = _golden_modules(g, "cpu")[0]  # prediction-network output computed on CPU, as in the golden run
        out_pn = cpu_pred.dec_lin(cpu_pred.dec(cpu_pred.emb(torch.zeros(tn.shape[0], 1, dtype=torch.int32)))[0]).to(d)
        for t in range(g["first_frames_logp"].shape[0]):
            lp = step(tn[:, t, :].unsqueeze(1).unsqueeze(1), out_pn.unsqueeze(1))
            assert lp.shape == (tn.shape[0], 1, 1, head.weight.shape[0])
            np.testing.assert_allclose(lp.squeeze(1).squeeze(1).cpu().numpy(), g["first_frames_logp"][t], atol=2e-5)


@pytest.mark.gpu
def test_patch_searcher_swaps_only_supported_chains():
    d = torch.device("cuda:0")
    H, V = 64, 30
    head = torch.nn.Linear(H, V).to(d)
    sb_like_head = types.SimpleNamespace(w=head)  # SpeechBrain's Linear keeps nn.Linear in .w (linear.py:61)
    searcher = types.SimpleNamespace(tjoint=tsasr_b200.Transducer_joint(joint="sum"), classifier_network=[sb_like_head],
                                     softmax=torch.nn.LogSoftmax(dim=-1))
    assert decode.patch_searcher(searcher)
    lp = searcher._joint_forward_step(torch.randn(4, 1, 1, H, device=d), torch.randn(4, 1, 1, H, device=d))
    assert lp.shape == (4, 1, 1, V) and torch.allclose(lp.exp().sum(-1), torch.ones(4, 1, 1, device=d), atol=1e-5)
    lp1 = searcher._joint_forward_step(torch.randn(1, 1, 1, H, device=d), torch.randn(1, 1, 1, H, device=d))  # beam search shape
    assert lp1.shape == (1, 1, 1, V)
    concat = types.SimpleNamespace(tjoint=tsasr_b200.Transducer_joint(joint="concat"), classifier_network=[head],
                                   softmax=torch.nn.LogSoftmax(dim=-1))
    assert not decode.patch_searcher(concat)
    two_layers = types.SimpleNamespace(tjoint=tsasr_b200.Transducer_joint(joint="sum"), classifier_network=[head, head],
                                       softmax=torch.nn.LogSoftmax(dim=-1))
    assert not decode.patch_searcher(two_layers)


def test_on_device_greedy_bookkeeping_reproduces_reference_searcher_cpu(golden):
    """CPU (no kernels involved): the device-side bookkeeping of ``greedy_decode_on_device`` -- masked select instead of
    per-utterance ``.item()`` branches -- gives exactly the reference searcher's hypotheses and score."""
    g = golden("greedy_decode")
    pred, head, tjoint, tn, hyps = _golden_modules(g, "cpu")
    step = eager_joint_step(tjoint, [head], torch.nn.LogSoftmax(dim=-1))
    got, score, a, b = decode.greedy_decode_on_device(tn, pred.layers(), step, blank_id=0)
    assert got == hyps and a is None and b is None
    np.testing.assert_allclose(float(score), float(g["mean_exp_score"]), rtol=1e-4)


@pytest.mark.gpu
def test_on_device_greedy_with_fused_step_vs_reference_golden(golden):
    """The whole greedy search on the GPU: fused joint step + on-device bookkeeping; ONE device->host copy at the end.
    Hypotheses bit-exact against the vector produced by the reference's own searcher class."""
    g = golden("greedy_decode")
    d = torch.device("cuda:0")
    pred, head, tjoint, tn, hyps = _golden_modules(g, d)
    torch.backends.cudnn.allow_tf32 = False
    searcher = types.SimpleNamespace(tjoint=tjoint, classifier_network=[head], softmax=torch.nn.LogSoftmax(dim=-1),
                                     decode_network_lst=pred.layers(), blank_id=0, beam_size=1)
    assert decode.patch_searcher(searcher, on_device_greedy=True)
    launches0 = tsasr_b200._lib.launch_count()
    got, score, _, _ = searcher.searcher(tn)
    assert tsasr_b200._lib.launch_count() - launches0 == 2 * tn.shape[1]
    assert got == hyps
    np.testing.assert_allclose(float(score), float(g["mean_exp_score"]), rtol=1e-4)
    # and it agrees with the per-frame-sync loop on a larger random problem (same modules, same fused step)
    gen = torch.Generator().manual_seed(5)
    tn2 = (1.5 * torch.randn(6, 120, tn.shape[2], generator=gen)).to(d)
    ref_h, ref_s = greedy_decode(tn2, pred.layers(), searcher._joint_forward_step)
    got_h, got_s, _, _ = searcher.searcher(tn2)
    assert got_h == ref_h
    np.testing.assert_allclose(float(got_s), np.exp(np.array(ref_s)).mean(), rtol=1e-4)


@pytest.mark.gpu
def test_cuda_graph_greedy_matches_the_plain_on_device_loop(golden):
    """One CUDA-graph launch per frame (frame index advanced on the device inside the graph): same hypotheses and score
    as the reference's searcher (golden vector) and as the plain on-device loop on a longer random problem."""
    g = golden("greedy_decode")
    d = torch.device("cuda:0")
    pred, head, tjoint, tn, hyps = _golden_modules(g, d)
    torch.backends.cudnn.allow_tf32 = False
    step = decode.fused_joint_forward_step(tjoint, [head], torch.nn.LogSoftmax(dim=-1))
    got, score, _, _ = decode.greedy_decode_cuda_graph(tn, pred.layers(), step, blank_id=0)
    assert got == hyps
    np.testing.assert_allclose(float(score), float(g["mean_exp_score"]), rtol=1e-4)
    gen = torch.Generator().manual_seed(6)
    tn2 = (1.5 * torch.randn(5, 90, tn.shape[2], generator=gen)).to(d)
    want_h, want_s, _, _ = decode.greedy_decode_on_device(tn2, pred.layers(), step, blank_id=0)
    got_h, got_s, _, _ = decode.greedy_decode_cuda_graph(tn2, pred.layers(), step, blank_id=0)
    assert got_h == want_h
    np.testing.assert_allclose(float(got_s), float(want_s), rtol=1e-5)


# ---- beam search: all utterances searched concurrently (tsasr_b200.decode.beam_search_batched) ----
def _beam_golden(g, gb, device):
    pred, head, tjoint, tn, _ = _golden_modules(g, device)
    with torch.no_grad():
        head.weight.copy_(torch.tensor(gb["W"]).to(device))
        head.bias.copy_(torch.tensor(gb["b"]).to(device))
    beam_size, nbest = (int(x) for x in gb["cfg"])
    state_beam, expand_beam = (float(x) for x in gb["beams"])
    B = tn.shape[0]
    hyps, o = [], 0
    for n in gb["hyp_lens"]:
        hyps.append([int(x) for x in gb["hyp_flat"][o:o + int(n)]])
        o += int(n)
    nbest_ref, scores_ref, o = [], [], 0
    for n in gb["hyp_counts"]:
        nbest_ref.append(hyps[o:o + int(n)])
        scores_ref.append([float(x) for x in gb["scores"][o:o + int(n)]])
        o += int(n)
    assert len(nbest_ref) == B
    return pred, head, tjoint, tn, dict(beam_size=beam_size, nbest=nbest, state_beam=state_beam, expand_beam=expand_beam), nbest_ref, scores_ref


def test_batched_beam_search_reproduces_reference_searcher_cpu(golden):
    """CPU (no kernels involved): the concurrent search -- one coroutine per utterance, one batched network evaluation
    per round -- gives exactly the n-best lists and scores of the reference's sequential searcher class
    (tests/golden/beam_decode.npz, made by oracle/make_golden_beam.py with the REAL TransducerBeamSearcher)."""
    g, gb = golden("greedy_decode"), golden("beam_decode")
    pred, head, tjoint, tn, cfg, nbest_ref, scores_ref = _beam_golden(g, gb, "cpu")
    step = eager_joint_step(tjoint, [head], torch.nn.LogSoftmax(dim=-1))
    best, score, nbest, nbest_scores = decode.beam_search_batched(tn, pred.layers(), step, blank_id=0, **cfg)
    assert nbest == nbest_ref and best == [n[0] for n in nbest_ref]
    for got, want in zip(nbest_scores, scores_ref):
        np.testing.assert_allclose(np.array(got, dtype=np.float64), want, rtol=1e-5)
    np.testing.assert_allclose(float(score), float(gb["mean_exp_score"]), rtol=1e-5)


@pytest.mark.gpu
def test_batched_beam_search_with_fused_step_vs_reference_golden(golden):
    """The same on the GPU with the fused joint step, through ``patch_searcher`` (beam_size > 1): n-best label sequences
    bit-exact against the reference's searcher, one device->host copy per expansion round."""
    g, gb = golden("greedy_decode"), golden("beam_decode")
    d = torch.device("cuda:0")
    pred, head, tjoint, tn, cfg, nbest_ref, scores_ref = _beam_golden(g, gb, d)
    torch.backends.cudnn.allow_tf32 = False
    searcher = types.SimpleNamespace(tjoint=tjoint, classifier_network=[head], softmax=torch.nn.LogSoftmax(dim=-1),
                                     decode_network_lst=pred.layers(), blank_id=0, lm_weight=0.0, **cfg)
    assert decode.patch_searcher(searcher, on_device_greedy=True)
    launches0 = tsasr_b200._lib.launch_count()
    best, score, nbest, nbest_scores = searcher.searcher(tn)
    assert tsasr_b200._lib.launch_count() > launches0  # the fused joint step ran
    assert nbest == nbest_ref
    for got, want in zip(nbest_scores, scores_ref):
        np.testing.assert_allclose(np.array(got, dtype=np.float64), want, rtol=1e-4)
    np.testing.assert_allclose(float(score), float(gb["mean_exp_score"]), rtol=1e-4)
