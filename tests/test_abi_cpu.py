"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/tsasr_b200.h declares (no compute calls without a GPU), and the host logic that needs no
kernels (length conversion, handle mechanics, error behaviour)."""
import ctypes
import os
import re

import pytest
import torch

import tsasr_b200
from tsasr_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tsasr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsasr_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    _build.build()
    lib = ctypes.CDLL(_build.SO_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tsasr_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert _lib.load().tsasr_abi_version() == _lib.ABI_VERSION == 4
    assert _lib.load().tsasr_lattice_elems(2, 5, 3) == 2 * 7 * 3


def test_no_cpu_path():
    logits = torch.randn(1, 3, 2, 5)
    with pytest.raises(ValueError, match="CUDA"):
        tsasr_b200.rnnt_loss(logits, torch.ones(1, 1, dtype=torch.int32), torch.tensor([3], dtype=torch.int32),
                             torch.tensor([1], dtype=torch.int32), blank=0, check_lengths=False)
    with pytest.raises(ValueError, match="cuda"):  # transducer_loss.py:348-351
        tsasr_b200.TransducerLoss()(logits, torch.ones(1, 1, dtype=torch.int32), torch.tensor([3], dtype=torch.int32),
                                    torch.tensor([1], dtype=torch.int32))


def test_error_conventions_match_torchaudio():
    from torchaudio.functional import rnnt_loss as ta

    logits = torch.randn(1, 3, 2, 5)
    good = dict(targets=torch.ones(1, 1, dtype=torch.int32), logit_lengths=torch.tensor([3], dtype=torch.int32),
                target_lengths=torch.tensor([1], dtype=torch.int32))
    for bad in (dict(good, targets=good["targets"].long()), dict(good, logit_lengths=good["logit_lengths"].long()),
                dict(good, target_lengths=good["target_lengths"].long())):
        with pytest.raises(RuntimeError):
            ta(logits, **bad)
        with pytest.raises(RuntimeError):
            tsasr_b200.rnnt_loss(logits, **bad)
    for fn in (ta, tsasr_b200.rnnt_loss):
        with pytest.raises(ValueError):
            fn(logits, reduction="batchmean", **good)
        with pytest.raises(RuntimeError):
            fn(logits, blank=7, **good)


def test_transducer_joint_module_surface():
    j = tsasr_b200.Transducer_joint()
    assert isinstance(j.nonlinearity, torch.nn.LeakyReLU)
    assert list(j.state_dict().keys()) == [] and list(j.buffers()) == []  # strict checkpoint loading
    # CPU / decode-shaped inputs use the reference's eager math
    tn, pn = torch.rand(8, 200, 1, 40), torch.rand(8, 1, 12, 40)
    out = j(tn, pn)
    assert torch.equal(out, torch.nn.functional.leaky_relu(tn + pn))
    lin = torch.nn.Linear(80, 80)
    jc = tsasr_b200.Transducer_joint(lin, joint="concat")
    assert jc(tn, pn).shape == torch.Size([8, 200, 12, 80])  # the reference's doctest (transducer_joint.py:27-38)
    with pytest.raises(ValueError):
        j(torch.rand(3, 4), torch.rand(3))


def test_joint_handle_defers_through_stock_linear_and_materialises_on_foreign_ops():
    enc, dec = torch.randn(2, 6, 1, 64), torch.randn(2, 1, 5, 64)
    act = torch.nn.LeakyReLU()
    h = tsasr_b200.JointHandle(enc, dec, act, 0, 0.01)
    assert h.shape == (2, 6, 5, 64) and h.ndim == 4 and h.dim() == 4 and h.device == enc.device
    head = torch.nn.Linear(64, 11)
    out = head(h)  # nn.Linear.forward -> F.linear(handle, W, b), exactly what SB/nnet/linear.py:74 issues
    assert isinstance(out, tsasr_b200.JointHandle) and out.has_head and out.shape == (2, 6, 5, 11)
    assert out._weight is head.weight and out._bias is head.bias
    ref = head(act(enc + dec))
    assert torch.allclose(out.materialize(), ref)
    foreign = out * 1.0
    assert type(foreign) is torch.Tensor and torch.allclose(foreign, ref)
    assert torch.allclose(out.log_softmax(-1), ref.log_softmax(-1))


def test_relative_length_conversion_bit_exact(golden):
    g = golden("half_rounding_torchaudio")
    rel = torch.tensor(g["input_rel"])
    assert torch.equal((rel * 8).round().int(), torch.tensor(g["input_abs"]))


@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_numba_branch_reduction_node_matches_the_reference_scaling(reduction):
    """``_NumbaReduce`` (the value / gradient scale of the reference's Numba branch on top of the fused per-utterance
    costs): value reduce_b(cost_b / T_b) (SB/nnet/loss/transducer_loss.py:104-106,280-287), gradient w.r.t. cost_b =
    grad_output broadcast -- the reduction and the division are NOT differentiated (:289-293)."""
    from tsasr_b200.functional import _NumbaReduce

    costs = torch.tensor([10.0, 20.0, 36.0], requires_grad=True)
    T = torch.tensor([5, 4, 9], dtype=torch.int32)
    out = _NumbaReduce.apply(costs, T, reduction)
    per_utt = torch.tensor([2.0, 5.0, 4.0])
    want = {"mean": per_utt.mean(), "sum": per_utt.sum(), "none": per_utt}[reduction]
    assert torch.allclose(out, want)
    gout = torch.tensor([1.0, 2.0, 3.0]) if reduction == "none" else torch.tensor(0.5)
    out.backward(gout)
    assert torch.equal(costs.grad, gout.expand(3))
    with pytest.raises(Exception, match="Unexpected reduction"):
        _NumbaReduce.apply(costs, T, "max")


def test_install_into_speechbrain_patches_the_three_import_names(monkeypatch):
    """``install_into_speechbrain`` swaps exactly the names the recipe's yaml resolves (conformer-t_scratch.yaml:191-193,
    262-264 -> speechbrain.nnet.transducer.transducer_joint.Transducer_joint, speechbrain.nnet.losses.transducer_loss)
    plus the Numba-branch classes; checked on stub modules (speechbrain itself is not importable on the test box)."""
    import sys
    import types

    names = ["speechbrain", "speechbrain.nnet", "speechbrain.nnet.losses", "speechbrain.nnet.transducer",
             "speechbrain.nnet.transducer.transducer_joint", "speechbrain.nnet.loss", "speechbrain.nnet.loss.transducer_loss"]
    mods = {n: types.ModuleType(n) for n in names}
    for n, m in mods.items():
        monkeypatch.setitem(sys.modules, n, m)
        if "." in n:
            setattr(mods[n.rsplit(".", 1)[0]], n.rsplit(".", 1)[1], m)
    mods["speechbrain.nnet.losses"].transducer_loss = object()
    mods["speechbrain.nnet.transducer.transducer_joint"].Transducer_joint = object()
    tsasr_b200.install_into_speechbrain()
    assert mods["speechbrain.nnet.losses"].transducer_loss is tsasr_b200.transducer_loss
    assert mods["speechbrain.nnet.transducer.transducer_joint"].Transducer_joint is tsasr_b200.Transducer_joint
    assert mods["speechbrain.nnet.loss.transducer_loss"].TransducerLoss is tsasr_b200.TransducerLoss
    assert mods["speechbrain.nnet.loss.transducer_loss"].Transducer is tsasr_b200.Transducer


def test_deferred_finite_check_keeps_the_reference_bookkeeping():
    """tsasr_b200.monitor: CPU tensors fall through to the original check; the deferred path (exercised with a stub
    that pretends to be a CUDA loss is covered on the GPU by tests/test_insitu_gpu.py)."""
    import torch

    import tsasr_b200

    class Brain:
        def __init__(self):
            self.calls = []

        def check_gradients(self, loss):
            self.calls.append(float(loss))
            return bool(loss.isfinite())

    b = Brain()
    chk = tsasr_b200.monitor.install(b)
    assert b.check_gradients(torch.tensor(1.0)) is True and b.check_gradients(torch.tensor(float("nan"))) is False
    assert len(b.calls) == 2 and chk.deferred_steps == 0
    tsasr_b200.monitor.uninstall(b)
    assert b.check_gradients.__func__ is Brain.check_gradients


def test_linear_dropin_surface_on_cpu():
    """tsasr_b200.Linear mirrors SB/nnet/linear.py:18-76: same constructor rules and parameter names; on CPU tensors it is
    the reference's own op (the CUDA kernels have no CPU path)."""
    with pytest.raises(ValueError, match="Expected one of input_shape or input_size"):
        tsasr_b200.Linear(4)
    lin = tsasr_b200.Linear(6, input_shape=[None, None, 5, 3], combine_dims=True)
    assert lin.w.in_features == 15 and [n for n, _ in lin.named_parameters()] == ["w.weight", "w.bias"]
    x = torch.randn(2, 7, 5, 3)
    want = torch.nn.functional.linear(x.reshape(2, 7, 15), lin.w.weight, lin.w.bias)
    assert torch.equal(lin(x), want)
    assert tsasr_b200.linear.bf16_twin(want) is None  # CPU tensors never carry a bf16 copy
    sd = torch.nn.Linear(15, 6).state_dict()
    lin.w.load_state_dict(sd)  # reference checkpoints (w.weight / w.bias) load unchanged


def test_projection_oracle_against_torch():
    import numpy as np
    from oracle import projection

    g = torch.Generator().manual_seed(0)
    x = torch.randn(9, 20, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(7, 20, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(7, generator=g, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.linear(x, w, b)
    dy = torch.randn(9, 7, generator=g, dtype=torch.float64)
    y.backward(dy)
    np.testing.assert_allclose(projection.linear_fwd(x.detach().numpy(), w.detach().numpy(), b.detach().numpy()), y.detach().numpy(), rtol=1e-12)
    dx, dw, db = projection.linear_bwd(dy.numpy(), x.detach().numpy(), w.detach().numpy())
    np.testing.assert_allclose(dx, x.grad.numpy(), rtol=1e-12)
    np.testing.assert_allclose(dw, w.grad.numpy(), rtol=1e-12)
    np.testing.assert_allclose(db, b.grad.numpy(), rtol=1e-12)


def test_adopt_modules_swaps_in_place_and_shares_parameters():
    """tsasr_b200.adopt_modules on modules built from the REFERENCE's class shapes (same class names and attributes as
    SB/nnet/linear.py, embedding.py, RNN.py): the drop-ins take over the very same parameter tensors."""
    class Linear(torch.nn.Module):          # stand-ins with the reference's names / attributes (speechbrain is not importable on CPU CI)
        def __init__(self):
            super().__init__()
            self.combine_dims = False
            self.w = torch.nn.Linear(6, 4)

    class Embedding(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.num_embeddings, self.consider_as_one_hot, self.embedding_dim, self.blank_id = 7, True, 6, 0
            self.Embedding = torch.nn.Embedding(7, 6, padding_idx=0)

    class LSTM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.reshape = False
            self.rnn = torch.nn.LSTM(6, 8, batch_first=True)

    mods = {"encoder_proj": Linear(), "decoder_proj": Linear(), "embedding": Embedding(), "decoder": LSTM(), "joiner": torch.nn.Identity()}
    old = dict(mods)
    assert tsasr_b200.adopt_modules(mods) == ["encoder_proj", "decoder_proj", "embedding", "decoder"]
    assert isinstance(mods["encoder_proj"], tsasr_b200.Linear) and mods["encoder_proj"].w is old["encoder_proj"].w
    assert isinstance(mods["embedding"], tsasr_b200.Embedding) and mods["embedding"].Embedding is old["embedding"].Embedding
    assert isinstance(mods["decoder"], tsasr_b200.LSTM) and mods["decoder"].rnn is old["decoder"].rnn
    assert mods["joiner"] is old["joiner"]
    assert [n for n, _ in mods["decoder"].named_parameters()] == [n for n, _ in old["decoder"].named_parameters()]
    # CPU tensors: the adopted modules compute what the originals compute
    tok = torch.tensor([[0, 3, 5]])
    x = mods["embedding"](tok)
    y, _ = mods["decoder"](x)
    assert torch.equal(y, old["decoder"].rnn(old["embedding"].Embedding(tok))[0])
    assert tsasr_b200.adopt_modules(mods) == []  # idempotent
