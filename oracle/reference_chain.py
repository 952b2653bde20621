"""The reference's own call path, restated with torch ops on CPU tensors (TEST ORACLE ONLY).

This is "Route 1" of SURVEY.md section 8c -- what ``compute_forward``/``compute_objectives``
execute (train_librispeechmix_scratch.py:122-135,150-160) -- using the third-party
``torchaudio.functional.rnnt_loss`` exactly the way
vendor/speechbrain/speechbrain/nnet/losses.py:58-59,72-79 reaches it.  It never imports
/root/reference (absent on the GPU box); torch and torchaudio ship in the image.
"""
import torch


def reference_transducer_loss(logits, targets, input_lens, target_lens, blank_index, reduction="mean"):
    """SB/nnet/losses.py:29-79 with use_torchaudio=True, verbatim semantics."""
    from torchaudio.functional import rnnt_loss

    input_lens = (input_lens * logits.shape[1]).round().int()  # losses.py:58
    target_lens = (target_lens * targets.shape[1]).round().int()  # losses.py:59
    return rnnt_loss(logits, targets.int(), input_lens, target_lens, blank=blank_index, reduction=reduction)


def reference_rnnt_abs(logits, targets, logit_lengths, target_lengths, blank=0, reduction="none"):
    """torchaudio rnnt_loss with absolute int32 lengths (the op under losses.py:72-79)."""
    from torchaudio.functional import rnnt_loss

    return rnnt_loss(logits, targets.int(), logit_lengths.int(), target_lengths.int(), blank=blank, reduction=reduction)


_ACTS = {
    "leaky_relu": lambda p: torch.nn.LeakyReLU(p),
    "relu": lambda p: torch.nn.ReLU(),
    "tanh": lambda p: torch.nn.Tanh(),
    "identity": lambda p: torch.nn.Identity(),
}


def reference_joint_logits(enc, dec, W, bias, act="leaky_relu", act_param=0.01, round_bf16=False):
    """joiner(enc[..., None, :], dec[:, None, ...]) then transducer_head.

    SB/nnet/transducer/transducer_joint.py:73-74,95 (joint="sum") and SB/nnet/linear.py:74.
    ``round_bf16`` inserts the single rounding the fused kernel applies to the joint activations
    (operands enc/dec/W are expected to be bf16-representable already in that case).
    """
    joint = enc[..., None, :] + dec[:, None, ...]
    joint = _ACTS[act](act_param)(joint)
    if round_bf16:
        # straight-through rounding: value rounded, gradient identity (the kernel's act' uses the
        # rounded output, which has the same sign / nearly the same value)
        joint = joint + (joint.detach().to(torch.bfloat16).to(joint.dtype) - joint.detach())
    return torch.nn.functional.linear(joint, W, bias)


def reference_joint_loss_fwd_bwd(enc, dec, W, bias, targets, logit_lengths, target_lengths, blank=0,
                                 act="leaky_relu", act_param=0.01, round_bf16=False, reduction="none",
                                 dcost=None, num_threads=None, keep_intermediates=False):
    """Full CPU chain joint -> head -> rnnt_loss -> backward; returns dict of fp32 tensors.
    ``keep_intermediates`` adds "dlogits" [B,T,U,V] (the gradient the reference's autograd feeds into the head Linear,
    incl. dcost) and "joint" [B,T,U,H] (the activated, optionally bf16-rounded joint tensor)."""
    if num_threads:
        torch.set_num_threads(num_threads)
    enc = enc.detach().float().cpu().clone().requires_grad_()
    dec = dec.detach().float().cpu().clone().requires_grad_()
    W = W.detach().float().cpu().clone().requires_grad_()
    bias = bias.detach().float().cpu().clone().requires_grad_()
    if keep_intermediates:
        joint = _ACTS[act](act_param)(enc[..., None, :] + dec[:, None, ...])
        if round_bf16:
            joint = joint + (joint.detach().to(torch.bfloat16).to(joint.dtype) - joint.detach())
        logits = torch.nn.functional.linear(joint, W, bias)
        logits.retain_grad()
    else:
        joint = None
        logits = reference_joint_logits(enc, dec, W, bias, act, act_param, round_bf16)
    costs = reference_rnnt_abs(logits, targets.cpu(), logit_lengths.cpu(), target_lengths.cpu(), blank, "none")
    if dcost is None:
        dcost = torch.ones_like(costs)
    (costs * dcost.cpu().float()).sum().backward()
    loss = {"none": costs, "sum": costs.sum(), "mean": costs.mean()}[reduction]
    out = {"costs": costs.detach(), "loss": loss.detach(), "d_enc": enc.grad, "d_dec": dec.grad,
           "dW": W.grad, "db": bias.grad}
    if keep_intermediates:
        out["dlogits"] = logits.grad
        out["joint"] = joint.detach()
    return out
