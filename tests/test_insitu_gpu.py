"""BASELINE configs[4] in situ: one accumulation cycle of the REAL ``TSASR.fit_batch`` (train_librispeechmix_scratch.py:33-190,
SB/core.py:1032-1096) with every drop-in (``Transducer_joint`` / ``transducer_loss``, the projection ``Linear``s, the prediction
network's ``Embedding`` / ``LSTM``) against the same cycle with the stock SpeechBrain modules -- same initial weights, same synthetic batch, same dropout seeds.  Needs the reference install under
baseline/_ref (tools/install_reference.sh, travels to the GPU box); skipped when it is absent."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.timeout(600)
@pytest.mark.parametrize("V", [256, 29])  # 29: the recipe's own character vocabulary -> the narrow-vocabulary form of the joint kernel
def test_one_fit_batch_cycle_matches_the_stock_recipe(V):
    import insitu_step as ins

    if ins.find_reference() is None:
        pytest.skip("no reference install (baseline/_ref): run tools/install_reference.sh in the authoring container")
    import tsasr_b200

    sb, rec, ConformerEncoder = ins.import_reference()
    dev = torch.device("cuda:0")
    # 6 s of audio -> T = 151 frames, 30 labels, V = 256 / 29: the loss is a sum over ~180 lattice steps, so the bf16 rounding of
    # the drop-in's GEMM operands (the stock arm is fp32 throughout) stays well inside north_star's 1e-4 relative
    gaf, seed = 2, 1234
    batches = [ins.make_batch(sb, B=3, seconds=6.0, n_labels=30, V=V, seed=i, ragged=True) for i in range(2)]
    res = {}
    launches0 = tsasr_b200._lib.launch_count()
    for arm, dropin in (("stock", False), ("dropin", True)):
        brain = ins.make_brain(sb, rec, ConformerEncoder, V, dropin, dev, False, gaf, seed, dropout=0.1)
        losses, grads = ins.parity_cycle(brain, batches, gaf, seed)
        res[arm] = {"losses": losses, **grads}
        assert brain.optimizer_step == 1                       # the cycle ended with an optimizer step
        if dropin:
            assert isinstance(brain.modules.joiner, tsasr_b200.Transducer_joint)
            assert isinstance(getattr(brain.modules.decoder, "module", brain.modules.decoder), tsasr_b200.LSTM)
            assert isinstance(getattr(brain.modules.encoder_proj, "module", brain.modules.encoder_proj), tsasr_b200.Linear)
            assert tsasr_b200._lib.launch_count() > launches0  # the fused kernels ran inside fit_batch
        else:
            assert tsasr_b200._lib.launch_count() == launches0
        del brain
    par = ins.compare(res["stock"], res["dropin"])
    # loss: 1e-4 relative (north_star).  Gradients: the drop-in rounds enc_out, dec_out, W and dlogits to bf16 (the
    # stock arm is fp32 throughout), so they agree to bf16-operand accuracy, not to fp32 rounding
    assert par["loss_rel_err_max"] < 1e-4, par
    assert par["head_grad_max_err_over_max"] < 1e-2 and par["head_bias_grad_max_err_over_max"] < 1e-2, par
    assert par["enc_proj_grad_max_err_over_max"] < 2e-2, par
    assert par["dec_rnn_hh_grad_max_err_over_max"] < 2e-2 and par["dec_rnn_ih_grad_max_err_over_max"] < 2e-2, par
