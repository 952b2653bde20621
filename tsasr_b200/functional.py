"""Autograd functions and the functional API of the B200-native RNN-T loss.

``rnnt_loss`` mirrors ``torchaudio.functional.rnnt_loss`` -- the op the reference reaches at
vendor/speechbrain/speechbrain/nnet/losses.py:72-79 -- in signature, semantics and error behaviour
(same exception types for the same precondition violations).
"""
import threading

import torch

from . import _lib, ops


def _check_reduction(reduction):
    if reduction not in ("none", "mean", "sum"):
        raise ValueError('reduction should be one of "none", "mean", or "sum"')


MAX_LATTICE_WIDTH = 1024  # alpha_beta_kernel: one thread per lattice column (the reference's Numba kernels share this limit)


def _check_lattice_shape(U, V):
    """Limits of the sm_100a kernels, rejected where the call is made -- not three launches later inside the C library."""
    if U > MAX_LATTICE_WIDTH:
        raise NotImplementedError(
            f"tsasr_b200: lattice width U = max target length + 1 = {U} exceeds {MAX_LATTICE_WIDTH} (one DP thread per "
            "lattice column); split the utterance or shorten the targets")
    if V < 2:
        raise ValueError(f"tsasr_b200: the vocabulary must hold the blank and at least one label (got V = {V})")


def _reduce(costs, reduction):
    # torchaudio/functional/functional.py:1791-1794: reduction is applied OUTSIDE the autograd
    # Function, so "mean" back-propagates 1/B.
    if reduction == "mean":
        return costs.mean()
    if reduction == "sum":
        return costs.sum()
    return costs


def _validate_lengths(logit_lengths, target_lengths, T, U, n_targets):
    """The checks torchaudio performs on the host (one small device->host read, as the reference does)."""
    stats = torch.stack([logit_lengths.max(), target_lengths.max(), logit_lengths.min(), target_lengths.min()]).tolist()
    max_t, max_l, min_t, min_l = (int(v) for v in stats)
    if max_t != T:
        raise RuntimeError("input length mismatch")
    if max_l + 1 != U:
        raise RuntimeError("output length mismatch")
    if n_targets != max_l:
        raise RuntimeError("target length mismatch")
    if min_t < 1 or min_l < 0:
        raise RuntimeError("logit_lengths must be >= 1 and target_lengths >= 0")


class _DeferredLengthCheck:
    """The same host-side checks, read back on a side stream.  ``stats`` = int32 [max T_b, max labels, min T_b,
    min labels] produced on the caller's stream by ONE small kernel (``tsasr_prepare_lengths``) before the fused
    kernels are queued; ``ready`` is an event recorded right after it.  The copy to pinned memory runs on a second
    stream, so ``finish`` -- called after the fused kernels have been queued -- only waits for that small kernel
    while the GPU already works on the joint GEMM.  The fused kernels clamp every length to the padded lattice,
    so an invalid input cannot index out of bounds before ``finish`` raises (same exception types and messages
    as ``_validate_lengths``)."""

    _tls = threading.local()  # side stream + pinned landing buffer per (thread, device): calls from two threads never share one

    def __init__(self, stats, ready):
        dev = stats.device
        cache = self._tls.__dict__.setdefault("cache", {})
        if dev not in cache:
            cache[dev] = (torch.cuda.Stream(dev), torch.empty((4,), dtype=torch.int32).pin_memory())
        side, self.host = cache[dev]
        side.wait_event(ready)
        with torch.cuda.stream(side):
            self.host.copy_(stats, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record(side)
        stats.record_stream(side)

    def finish(self, T, U, n_targets):
        self.done.synchronize()
        max_t, max_l, min_t, min_l = (int(v) for v in self.host.tolist())
        if max_t != T:
            raise RuntimeError("input length mismatch")
        if max_l + 1 != U:
            raise RuntimeError("output length mismatch")
        if n_targets != max_l:
            raise RuntimeError("target length mismatch")
        if min_t < 1 or min_l < 0:
            raise RuntimeError("logit_lengths must be >= 1 and target_lengths >= 0")


def _prepare_lengths(logit_lengths, target_lengths, T, n_targets, relative):
    """-> (int32 logit_lengths, int32 target_lengths, stats [4] int32, event recorded after the kernel).

    relative=True: SpeechBrain relative lengths, converted exactly like SB/nnet/losses.py:58-59
    (fp32 product, round-half-to-even, int32) inside the one kernel that also computes the statistics."""
    dev = logit_lengths.device
    B = logit_lengths.shape[0]
    stats = torch.empty((4,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream(dev)
    if relative:
        rl = logit_lengths.to(torch.float32).contiguous()
        rt = target_lengths.to(torch.float32).contiguous()
        out = torch.empty((2, B), dtype=torch.int32, device=dev)
        ll, tl = out[0], out[1]
        _lib.check(lib.tsasr_prepare_lengths(rl.data_ptr(), rt.data_ptr(), None, None, B, int(T), int(n_targets),
                                             ll.data_ptr(), tl.data_ptr(), stats.data_ptr(), stream.cuda_stream))
    else:
        ll = logit_lengths.to(torch.int32).contiguous()
        tl = target_lengths.to(torch.int32).contiguous()
        _lib.check(lib.tsasr_prepare_lengths(None, None, ll.data_ptr(), tl.data_ptr(), B, int(T), int(n_targets),
                                             None, None, stats.data_ptr(), stream.cuda_stream))
    ready = torch.cuda.Event()
    ready.record(stream)
    return ll, tl, stats, ready


class RnntLossFromLogits(torch.autograd.Function):
    """costs[b] = -log P(y_b | x_b) from materialised logits; backward emits dense dlogits.

    Compat path: three HBM-bound kernels (log-sum-exp + gather, wavefront DP, gradient); the logits
    are read once in forward and once in backward, dlogits is written once.
    """

    @staticmethod
    def forward(ctx, logits, targets, logit_lengths, target_lengths, blank, clamp):
        B, T, U, V = logits.shape
        lat2, den = ops.logits_to_lattice(logits, targets, logit_lengths, target_lengths, blank)
        alpha, beta, cost, _, _ = ops.alpha_beta(lat2, logit_lengths, target_lengths, B, T, U)
        ctx.save_for_backward(logits, targets, logit_lengths, target_lengths, lat2, den, alpha, beta, cost)
        ctx.blank, ctx.clamp = blank, clamp
        return cost.to(logits.dtype)

    @staticmethod
    def backward(ctx, dcost):
        logits, targets, ll, tl, lat2, den, alpha, beta, cost = ctx.saved_tensors
        dcost = dcost.to(torch.float32).contiguous()
        dlogits = ops.logits_grad(logits, targets, ll, tl, ctx.blank, lat2, den, alpha, beta, cost, dcost, ctx.clamp)
        return dlogits, None, None, None, None, None


def rnnt_loss(logits, targets, logit_lengths, target_lengths, blank=-1, clamp=-1.0, reduction="mean",
              fused_log_softmax=True, check_lengths=True):
    """Drop-in for ``torchaudio.functional.rnnt_loss`` on CUDA tensors (absolute int32 lengths)."""
    _check_reduction(reduction)
    if logits.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise RuntimeError("logits must be float32, float16 or bfloat16")
    if targets.dtype != torch.int32:
        raise RuntimeError("targets must be int32 type")
    if logit_lengths.dtype != torch.int32:
        raise RuntimeError("logit_lengths must be int32 type")
    if target_lengths.dtype != torch.int32:
        raise RuntimeError("target_lengths must be int32 type")
    if logits.dim() != 4:
        raise RuntimeError("logits must be 4-D (batch, time, target, class)")
    if targets.dim() != 2:
        raise RuntimeError("targets must be 2-D (batch, max target length)")
    if not logits.is_contiguous():
        raise RuntimeError("logits must be contiguous")
    B, T, U, V = logits.shape
    if targets.shape[0] != B or logit_lengths.shape[0] != B or target_lengths.shape[0] != B:
        raise RuntimeError("batch dimension mismatch between logits, targets and lengths")
    if blank < 0:
        blank += V  # torchaudio accepts negative indices (default -1)
    if not 0 <= blank < V:
        raise RuntimeError("blank must be within [0, logits.shape[-1])")
    _check_lattice_shape(U, max(V, 2))  # V = 1 (blank only) is legal for torchaudio and for the compat kernels
    if check_lengths:
        _validate_lengths(logit_lengths, target_lengths, T, U, targets.shape[1])
    if not fused_log_softmax:
        logits = torch.nn.functional.log_softmax(logits, dim=-1)
    costs = RnntLossFromLogits.apply(logits, targets.contiguous(), logit_lengths.contiguous(),
                                     target_lengths.contiguous(), int(blank), float(clamp))
    return _reduce(costs, reduction)


class NumbaSemanticsTransducer(torch.autograd.Function):
    """``Transducer.apply(log_probs, labels, T, U, blank, reduction)`` --
    vendor/speechbrain/speechbrain/nnet/loss/transducer_loss.py:239-293, quirks included:
    value = reduce_b(-log P_b / T_b) with the reduction applied inside forward (:280-287), stored
    gradient w.r.t. the log-probs is the un-normalised occupation (:183-236), backward multiplies it by
    grad_output (:289-293)."""

    @staticmethod
    def forward(ctx, log_probs, labels, T, U, blank, reduction):
        log_probs = log_probs.detach()
        if log_probs.dtype != torch.float32:
            raise TypeError("log_probs must be float32 (the reference kernels are typed float32[:,:,:,:])")
        if labels.dtype != torch.int32 or T.dtype != torch.int32 or U.dtype != torch.int32:
            raise TypeError("labels, T and U must be int32 (the reference kernels are typed int32)")
        Bn, maxT, maxU, A = log_probs.shape
        _check_lattice_shape(maxU, max(A, 2))
        log_probs = log_probs.contiguous()
        lat2, _ = ops.logits_to_lattice(log_probs, labels, T, U, blank, normalized=True)
        alpha, beta, cost, _, _ = ops.alpha_beta(lat2, T, U, Bn, maxT, maxU)
        ones = torch.ones((Bn,), dtype=torch.float32, device=log_probs.device)
        ctx.grads = ops.logprobs_grad(tuple(log_probs.shape), labels, T, U, blank, lat2, alpha, beta, cost, ones)
        per_utt = cost / T.to(torch.float32)  # transducer_loss.py:104-106
        if reduction == "mean":
            return per_utt.mean()
        elif reduction == "sum":
            return per_utt.sum()
        elif reduction == "none":
            return per_utt
        else:
            raise Exception("Unexpected reduction {}".format(reduction))

    @staticmethod
    def backward(ctx, grad_output):
        grad_output = grad_output.view(-1, 1, 1, 1).to(ctx.grads)
        return ctx.grads.mul_(grad_output), None, None, None, None, None, None


def _operands_bf16(enc, dec, W):
    """bf16 copies of the three GEMM operands: one launch when all are contiguous fp32 (the recipe's case)."""
    ts = (enc, dec, W)
    if all(t.dtype == torch.float32 and t.is_contiguous() and t.numel() % 4 == 0 and t.data_ptr() % 16 == 0 for t in ts):
        # one allocation for the three copies: H % 64 == 0 here (FusedJointRnnt pads), so every slice starts 128-byte aligned
        n0, n1, n2 = (t.numel() for t in ts)
        buf = torch.empty((n0 + n1 + n2,), dtype=torch.bfloat16, device=enc.device)
        outs = (buf[:n0].view(enc.shape), buf[n0: n0 + n1].view(dec.shape), buf[n0 + n1:].view(W.shape))
        _lib.check(_lib.load().tsasr_cast_operands_bf16(
            enc.data_ptr(), enc.numel(), dec.data_ptr(), dec.numel(), W.data_ptr(), W.numel(),
            outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), torch.cuda.current_stream(enc.device).cuda_stream))
        return outs
    return tuple(t.to(torch.bfloat16).contiguous() for t in ts)


class FusedJointRnnt(torch.autograd.Function):
    """costs[b] of joint("sum") + activation + head Linear + RNN-T loss without the 4-D tensors.

    Differentiable inputs: enc_out [B,T,H], dec_out [B,U,H], W [V,H], bias [V]
    (train_librispeechmix_scratch.py:122,127,135).  Operands are rounded to bf16 once on entry; all
    accumulation is fp32 (TMEM), the lattice is fp32.
    """

    @staticmethod
    def forward(ctx, enc, dec, W, bias, targets, logit_lengths, target_lengths, blank, act_kind, act_param,
                max_chunk_cells, prune_log2_eps=None, clamp=-1.0):
        H = enc.shape[-1]
        enc_, dec_, W_ = enc.detach(), dec.detach(), W.detach()
        if H % 64:  # the kernels contract over whole 64-wide k-blocks: zero columns add nothing (act(0) = 0 for every
            pad = (0, 64 - H % 64)  # fused activation, and the padded W columns are zero anyway)
            enc_, dec_, W_ = (torch.nn.functional.pad(t, pad) for t in (enc_, dec_, W_))
        enc16, dec16, W16 = _operands_bf16(enc_, dec_, W_)
        b32 = bias.detach().to(torch.float32).contiguous()
        B, T, _ = enc16.shape
        U = dec16.shape[1]
        lat2, logz = ops.joint_fwd(enc16, dec16, W16, b32, targets, logit_lengths, target_lengths, blank, act_kind, act_param)
        alpha, beta, cost, _, _ = ops.alpha_beta(lat2, logit_lengths, target_lengths, B, T, U)
        ctx.save_for_backward(enc16, dec16, W16, b32, targets, logit_lengths, target_lengths, lat2, logz, alpha, beta, cost)
        ctx.cfg = (blank, act_kind, act_param, max_chunk_cells, prune_log2_eps, clamp)
        ctx.in_dtypes = (enc.dtype, dec.dtype, W.dtype, bias.dtype)
        ctx.H = H
        return cost

    @staticmethod
    def backward(ctx, dcost):
        enc16, dec16, W16, b32, targets, ll, tl, lat2, logz, alpha, beta, cost = ctx.saved_tensors
        blank, act_kind, act_param, max_chunk_cells, prune_log2_eps, clamp = ctx.cfg
        dcost = dcost.to(torch.float32).contiguous()
        d_enc, d_dec, dW, db = ops.joint_bwd(enc16, dec16, W16, b32, targets, ll, tl, blank, act_kind, act_param,
                                             lat2, logz, alpha, beta, cost, dcost, max_chunk_cells, prune_log2_eps, clamp)
        de, dd, dw, dbt = ctx.in_dtypes
        if d_enc.shape[-1] != ctx.H:  # drop the gradients of the zero padding
            d_enc, d_dec, dW = d_enc[..., :ctx.H].contiguous(), d_dec[..., :ctx.H].contiguous(), dW[:, :ctx.H].contiguous()
        return (d_enc.to(de), d_dec.to(dd), dW.to(dw), db.to(dbt), None, None, None, None, None, None, None, None, None)


class _NumbaReduce(torch.autograd.Function):
    """``reduce_b(cost_b / T_b)`` with the backward of the reference's Numba path: ``Transducer.backward``
    (SB/nnet/loss/transducer_loss.py:289-293) multiplies the stored per-utterance gradients by ``grad_output`` viewed as
    [-1,1,1,1] -- the reduction and the division by T_b (:104-106, :280-287) happen inside its forward and are NOT
    differentiated.  So d loss / d cost_b := grad_output (broadcast), whatever the reduction."""

    @staticmethod
    def forward(ctx, costs, T, reduction):
        ctx.n = costs.shape[0]
        per_utt = costs / T.to(torch.float32)
        if reduction == "mean":
            return per_utt.mean()
        elif reduction == "sum":
            return per_utt.sum()
        elif reduction == "none":
            return per_utt
        else:
            raise Exception("Unexpected reduction {}".format(reduction))

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.reshape(-1).to(torch.float32).expand(ctx.n), None, None


def fused_joint_rnnt_loss(enc_out, dec_out, weight, bias, targets, logit_lengths, target_lengths, blank=0,
                          activation="leaky_relu", act_param=0.01, reduction="mean", check_lengths=True,
                          max_chunk_cells=0, relative_lengths=False, prune_log2_eps=None, numba_semantics=False, clamp=-1.0):
    """Functional form of the fused path.  Lengths are absolute int32 counts, or -- with ``relative_lengths=True`` --
    SpeechBrain's relative floats, converted bit-exactly like SB/nnet/losses.py:58-59.
    ``prune_log2_eps``: backward tile pruning threshold (None: TSASR_PRUNE_LOG2_EPS or -30; >= 0: off), see
    include/tsasr_b200.h.
    ``clamp`` > 0: torchaudio's ``rnnt_loss(clamp=...)``: dlogits of the unit cost clamped to [-clamp, clamp] inside the
    gradient pass, before the upstream factor (1/B under "mean") multiplies them.
    ``numba_semantics``: value and gradient scale of the reference's Numba branch (``use_torchaudio=False`` /
    ``TransducerLoss``): ``reduce_b(-log P_b / T_b)``, gradient of ``sum_b -log P_b`` times the incoming grad_output."""
    if not numba_semantics:
        _check_reduction(reduction)
    if enc_out.dim() != 3 or dec_out.dim() != 3:
        raise ValueError("enc_out must be [B,T,H] and dec_out [B,U,H]")
    B, T, H = enc_out.shape
    U = dec_out.shape[1]
    V = weight.shape[0]
    if dec_out.shape[0] != B or dec_out.shape[2] != H or weight.shape[1] != H:
        raise ValueError("shape mismatch between enc_out, dec_out and weight")
    if not (enc_out.is_cuda and logit_lengths.is_cuda and target_lengths.is_cuda):
        raise ValueError("tsasr_b200 needs CUDA tensors; there is no CPU path")
    _check_lattice_shape(U, V)
    if bias is None:
        bias = torch.zeros((V,), dtype=torch.float32, device=enc_out.device)
    if blank < 0:
        blank += V
    if not 0 <= blank < V:
        raise RuntimeError("blank must be within [0, logits.shape[-1])")
    targets = targets.to(torch.int32).contiguous()
    with ops._on_device(enc_out.device):
        logit_lengths, target_lengths, stats, ready = _prepare_lengths(logit_lengths, target_lengths, T, targets.shape[1],
                                                                       relative_lengths)
        costs = FusedJointRnnt.apply(enc_out, dec_out, weight, bias, targets, logit_lengths, target_lengths, int(blank),
                                     _lib.ACT_CODES[activation], float(act_param), int(max_chunk_cells), prune_log2_eps, float(clamp))
        if check_lengths:
            # raises what torchaudio raises; the kernels above are already queued
            _DeferredLengthCheck(stats, ready).finish(T, U, targets.shape[1])
    if numba_semantics:
        return _NumbaReduce.apply(costs, logit_lengths, reduction)
    return _reduce(costs, reduction)
