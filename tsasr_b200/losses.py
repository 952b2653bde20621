"""Drop-ins for the SpeechBrain transducer-loss surface the ts-asr recipe uses.

* ``transducer_loss``  -- vendor/speechbrain/speechbrain/nnet/losses.py:29-87 (selected by
  hparams/LibriSpeechMix/conformer-t_scratch.yaml:262-264 with ``use_torchaudio: True``)
* ``Transducer`` / ``TransducerLoss`` -- vendor/speechbrain/speechbrain/nnet/loss/transducer_loss.py:239-351

Same names, argument meaning, value/gradient semantics (both of them) and
exception types; the arithmetic runs in libtsasr_b200.so.
"""
import torch
from torch.nn import Module

from . import functional as F
from .transducer_joint import JointHandle

Transducer = F.NumbaSemanticsTransducer
_CODE_TO_NAME = {0: "leaky_relu", 1: "relu", 2: "tanh", 3: "identity"}


def transducer_loss(logits, targets, input_lens, target_lens, blank_index, reduction="mean", use_torchaudio=True):
    """Transducer loss.

    Arguments
    ---------
    logits : torch.Tensor or JointHandle
        Predicted tensor of shape [batch, maxT, maxU, num_labels] -- or the deferred handle produced
        by ``tsasr_b200.Transducer_joint`` + the stock ``Linear`` head (fused path: the 4-D tensor is
        never materialised).
    targets : torch.Tensor
        Target tensor, without any blanks, of shape [batch, target_len].
    input_lens, target_lens : torch.Tensor
        RELATIVE lengths in (0, 1] (SpeechBrain convention).
    blank_index : int
    reduction : "mean" | "sum" | "none"
    use_torchaudio : bool
        True  -> torchaudio semantics (what the recipe runs): reduce_b(-log P_b), exact gradient.
        False -> the SpeechBrain Numba semantics: reduce_b(-log P_b / T_b), un-normalised gradient.
    """
    if isinstance(logits, JointHandle) and logits.has_head:
        # fused path, both semantics: torchaudio's (what the recipe runs) and the Numba branch's (losses.py:81-85)
        enc = logits._enc.squeeze(2)   # [B,T,1,H] -> [B,T,H]
        dec = logits._dec.squeeze(1)   # [B,1,U,H] -> [B,U,H]
        dev = enc.device
        fp32_lens = input_lens.dtype == torch.float32 and target_lens.dtype == torch.float32
        if not fp32_lens:  # unusual dtypes: the reference's own expression decides the rounding
            input_lens = (input_lens * logits.shape[1]).round().int()
            target_lens = (target_lens * targets.shape[1]).round().int()
        # fp32 relative lengths: the conversion of losses.py:58-59 runs, bit-exact, inside tsasr_prepare_lengths
        return F.fused_joint_rnnt_loss(
            enc, dec, logits._weight, logits._bias, targets.to(dev), input_lens.to(dev), target_lens.to(dev),
            blank=blank_index, activation=_CODE_TO_NAME[logits._act_code], act_param=logits._act_param,
            reduction=reduction, relative_lengths=fp32_lens, numba_semantics=not use_torchaudio,
            check_lengths=use_torchaudio)  # the Numba branch has no length preconditions (the kernels clamp)

    # integer length conversion, bit-exact with losses.py:58-59 (fp32 multiply, round-half-even, int32)
    input_lens = (input_lens * logits.shape[1]).round().int()
    target_lens = (target_lens * targets.shape[1]).round().int()

    if isinstance(logits, JointHandle):
        logits = logits.materialize()

    if use_torchaudio:
        return F.rnnt_loss(logits.contiguous(), targets.int(), input_lens, target_lens, blank=blank_index,
                           reduction=reduction)
    else:
        # Transducer.apply takes log-probs (losses.py:84); targets are NOT cast on this branch (:85)
        log_probs = logits.log_softmax(-1)
        return Transducer.apply(log_probs, targets, input_lens, target_lens, blank_index, reduction)


class TransducerLoss(Module):
    """``TransducerLoss(blank=0, reduction="mean").forward(logits, labels, T, U)`` with ABSOLUTE int32
    lengths; input tensors must be on a cuda device (transducer_loss.py:296-351)."""

    def __init__(self, blank=0, reduction="mean"):
        super(TransducerLoss, self).__init__()
        self.blank = blank
        self.reduction = reduction
        self.loss = Transducer.apply

    def forward(self, logits, labels, T, U):
        if all(t.is_cuda for t in (logits, labels, T, U)):
            if isinstance(logits, JointHandle) and logits.has_head:
                # fused path with the Numba branch's value / gradient scale; T, U are absolute int32 here
                return F.fused_joint_rnnt_loss(
                    logits._enc.squeeze(2), logits._dec.squeeze(1), logits._weight, logits._bias, labels, T, U,
                    blank=self.blank, activation=_CODE_TO_NAME[logits._act_code], act_param=logits._act_param,
                    reduction=self.reduction, numba_semantics=True, check_lengths=False)
            if isinstance(logits, JointHandle):
                logits = logits.materialize()
            log_probs = logits.log_softmax(-1)
            return self.loss(log_probs, labels, T, U, self.blank, self.reduction)
        else:
            raise ValueError(
                f"Found inputs tensors to be on {[logits.device, labels.device, T.device, U.device]} while needed to be on a 'cuda' device to use the transducer loss."
            )
