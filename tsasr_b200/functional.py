"""Autograd functions and the functional API of the B200-native RNN-T loss.

``rnnt_loss`` mirrors ``torchaudio.functional.rnnt_loss`` -- the op the reference reaches at
vendor/speechbrain/speechbrain/nnet/losses.py:72-79 -- in signature, semantics and error behaviour
(same exception types for the same precondition violations).
"""
import threading
import time

import torch

from . import _lib, ops


def _check_reduction(reduction):
    if reduction not in ("none", "mean", "sum"):
        raise ValueError('reduction should be one of "none", "mean", or "sum"')


MAX_LATTICE_WIDTH = 8192  # kMaxLatticeWidth (csrc/capi.cu): one DP thread per column up to 1024, several columns per thread beyond


def _check_lattice_shape(U, V):
    """Limits of the sm_100a kernels, rejected where the call is made -- not three launches later inside the C library."""
    if U > MAX_LATTICE_WIDTH:
        raise NotImplementedError(
            f"tsasr_b200: lattice width U = max target length + 1 = {U} exceeds {MAX_LATTICE_WIDTH} (the DP keeps two "
            "anti-diagonals in shared memory); split the utterance or shorten the targets")
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


class _StatsRing:
    """Landing slots (mapped pinned host memory) for the length statistics of the fused forward: the preparation kernel
    writes {max T_b, max labels, min T_b, min labels}, fences, then writes a sequence tag; the host polls the tag.  No
    stream, event or copy is involved, so torchaudio's argument checks cost a few microseconds of host time.  One ring per
    (thread, device); 64 slots, so a check may lag many calls behind without being overwritten."""

    _tls = threading.local()
    SLOTS = 64

    def __init__(self):
        self.buf = torch.zeros((self.SLOTS, 8), dtype=torch.int32).pin_memory()
        self.view = self.buf.numpy()
        self.seq = 0

    @classmethod
    def get(cls, dev):
        rings = cls._tls.__dict__.setdefault("rings", {})
        ring = rings.get(dev)
        if ring is None:
            ring = rings[dev] = cls()
        return ring

    def next_slot(self):
        self.seq = self.seq % 0x3FFFFFFF + 1
        slot = self.seq % self.SLOTS
        self.view[slot, 4] = 0
        return slot, self.seq, self.buf.data_ptr() + 32 * slot

    def wait(self, slot, seq, dev):
        """-> (max T_b, max labels, min T_b, min labels) once the kernel has delivered them."""
        row = self.view[slot]
        if row[4] != seq:
            deadline = time.perf_counter() + 20.0
            while row[4] != seq:
                if time.perf_counter() > deadline:  # never expected: fall back to a full synchronisation
                    torch.cuda.current_stream(dev).synchronize()
                    if row[4] != seq:
                        raise RuntimeError("tsasr_b200: the length statistics of the fused forward never arrived")
        return int(row[0]), int(row[1]), int(row[2]), int(row[3])


_layout_cache = {}


def _fwd_layout(B, T, U, H, V):
    key = (B, T, U, H, V)
    off = _layout_cache.get(key)
    if off is None:
        import ctypes

        arr = (ctypes.c_size_t * 8)()
        _lib.check(_lib.load().tsasr_joint_loss_fwd_layout(B, T, U, H, V, ctypes.cast(arr, ctypes.c_void_p)))
        off = _layout_cache[key] = tuple(int(x) for x in arr)
    return off


def _raise_length_errors(stats, T, U, n_targets):
    max_t, max_l, min_t, min_l = stats
    if max_t != T:
        raise RuntimeError("input length mismatch")
    if max_l + 1 != U:
        raise RuntimeError("output length mismatch")
    if n_targets != max_l:
        raise RuntimeError("target length mismatch")
    if min_t < 1 or min_l < 0:
        raise RuntimeError("logit_lengths must be >= 1 and target_lengths >= 0")


class FusedJointRnnt(torch.autograd.Function):
    """costs[b] of joint("sum") + activation + head Linear + RNN-T loss without the 4-D tensors.

    Differentiable inputs: enc_out [B,T,H], dec_out [B,U,H], W [V,H], bias [V]
    (train_librispeechmix_scratch.py:122,127,135).  Operands are rounded to bf16 once on entry; all
    accumulation is fp32 (TMEM), the lattice is fp32.

    The whole forward -- input preparation (bf16 operand copies, int32 targets, the integer length conversion of
    SB/nnet/losses.py:58-59 with its statistics), the joint GEMM with its online log-softmax and the alpha/beta DP -- is ONE
    call into the library (``tsasr_joint_loss_fwd``): the host time in front of the first GEMM launch is GPU idle time in
    a training loop that reads its loss every step (SB/core.py:1096).
    Returns (costs [B], int32 logit_lengths [B], int32 target_lengths [B]); only the first output is differentiable.
    """

    @staticmethod
    def forward(ctx, enc, dec, W, bias, targets, logit_lengths, target_lengths, relative_lengths, blank, act_kind, act_param,
                max_chunk_cells, prune_log2_eps=None, clamp=-1.0, check_lengths=True):
        dev = enc.device
        B, T, H = enc.shape
        U, V = dec.shape[1], W.shape[0]
        lib = _lib.load()
        ops_ = (enc.detach(), dec.detach(), W.detach())
        if H % 64 == 0 and all(t.dtype == torch.float32 and t.is_contiguous() and t.data_ptr() % 16 == 0 for t in ops_):
            code = _lib.F32        # the recipe's case: converted inside the preparation kernel
        elif H % 64 == 0 and all(t.dtype == torch.bfloat16 and t.is_contiguous() and t.data_ptr() % 16 == 0 for t in ops_):
            code = _lib.BF16       # projections that already emit bf16 (autocast): used as they are
        else:
            # fp16 / mixed dtypes / strided views / H not a multiple of 64: one torch pass.  The kernels contract over whole
            # 64-wide k-blocks; zero columns add nothing (act(0) = 0 for every fused activation, padded W columns are zero)
            pad = (0, (-H) % 64)
            ops_ = tuple(torch.nn.functional.pad(t, pad).to(torch.bfloat16).contiguous() if pad[1] else t.to(torch.bfloat16).contiguous()
                         for t in ops_)
            code = _lib.BF16
        Hp = ops_[0].shape[-1]
        # projections that ran through tsasr_linear_fwd (tsasr_b200.Linear) left the bf16 rounding of their output behind:
        # it IS the operand image, so the preparation kernel skips that operand
        tw_enc = tw_dec = None
        if code == _lib.F32:
            from .linear import bf16_twin

            tw_enc, tw_dec = bf16_twin(enc, consume=True), bf16_twin(dec, consume=True)
        b32 = bias.detach()
        if b32.dtype != torch.float32 or not b32.is_contiguous():
            b32 = b32.to(torch.float32).contiguous()
        tg = targets
        if tg.dtype not in (torch.int32, torch.int64) or not tg.is_contiguous():
            tg = tg.to(torch.int32).contiguous()
        n_targets = tg.shape[1]
        if relative_lengths:
            ll_in = logit_lengths if logit_lengths.dtype == torch.float32 and logit_lengths.is_contiguous() else logit_lengths.to(torch.float32).contiguous()
            tl_in = target_lengths if target_lengths.dtype == torch.float32 and target_lengths.is_contiguous() else target_lengths.to(torch.float32).contiguous()
            len_args = (ll_in.data_ptr(), tl_in.data_ptr(), None, None)
        else:
            ll_in = logit_lengths if logit_lengths.dtype == torch.int32 and logit_lengths.is_contiguous() else logit_lengths.to(torch.int32).contiguous()
            tl_in = target_lengths if target_lengths.dtype == torch.int32 and target_lengths.is_contiguous() else target_lengths.to(torch.int32).contiguous()
            len_args = (None, None, ll_in.data_ptr(), tl_in.data_ptr())
        off = _fwd_layout(B, T, U, Hp, V)
        scratch = torch.empty((off[7],), dtype=torch.uint8, device=dev)
        n = ops.lattice_elems(B, T, U)
        out = torch.empty((5 * n + 3 * B,), dtype=torch.float32, device=dev)  # lat2 | logz | alpha | beta | cost, ll_a, ll_b
        lat2, logz, alpha, beta, cost3 = out[: 2 * n].view(n, 2), out[2 * n: 3 * n], out[3 * n: 4 * n], out[4 * n: 5 * n], out[5 * n:]
        ring = _StatsRing.get(dev) if check_lengths else None
        slot, seq, slot_ptr = ring.next_slot() if ring is not None else (0, 0, None)
        with ops._on_device(dev):
            _lib.check(lib.tsasr_joint_loss_fwd(
                ops_[0].data_ptr(), ops_[1].data_ptr(), ops_[2].data_ptr(), code, b32.data_ptr(), tg.data_ptr() if tg.numel() else None,
                1 if tg.dtype == torch.int64 else 0, *len_args, B, T, U, Hp, V, int(blank), int(act_kind), float(act_param),
                scratch.data_ptr(), scratch.numel(), slot_ptr, seq, lat2.data_ptr(), logz.data_ptr(), alpha.data_ptr(), beta.data_ptr(),
                cost3.data_ptr(), tw_enc.data_ptr() if tw_enc is not None else None, tw_dec.data_ptr() if tw_dec is not None else None,
                torch.cuda.current_stream(dev).cuda_stream))
        if code == _lib.F32:
            enc16 = tw_enc if tw_enc is not None else scratch[off[0]: off[0] + 2 * B * T * Hp].view(torch.bfloat16).view(B, T, Hp)
            dec16 = tw_dec if tw_dec is not None else scratch[off[1]: off[1] + 2 * B * U * Hp].view(torch.bfloat16).view(B, U, Hp)
            W16 = scratch[off[2]: off[2] + 2 * V * Hp].view(torch.bfloat16).view(V, Hp)
        else:
            enc16, dec16, W16 = ops_
        tg32 = scratch[off[3]: off[3] + 4 * B * n_targets].view(torch.int32).view(B, n_targets) if tg.dtype == torch.int64 else tg
        ll = scratch[off[4]: off[4] + 4 * B].view(torch.int32)
        tl = scratch[off[5]: off[5] + 4 * B].view(torch.int32)
        cost = cost3[:B]
        ctx.save_for_backward(enc16, dec16, W16, b32, tg32, ll, tl, lat2, logz, alpha, beta, cost)
        ctx.cfg = (blank, act_kind, act_param, max_chunk_cells, prune_log2_eps, clamp)
        ctx.in_dtypes = (enc.dtype, dec.dtype, W.dtype, bias.dtype)
        ctx.H = H
        ctx.mark_non_differentiable(ll, tl)
        if ring is not None:
            # torchaudio's argument checks: same exception types and messages, raised after the (length-clamping) kernels
            # have been queued; the statistics arrive through mapped pinned memory, no synchronisation of the stream
            _raise_length_errors(ring.wait(slot, seq, dev), T, U, n_targets)
        return cost, ll, tl

    @staticmethod
    def backward(ctx, dcost, _dll=None, _dtl=None):
        enc16, dec16, W16, b32, targets, ll, tl, lat2, logz, alpha, beta, cost = ctx.saved_tensors
        blank, act_kind, act_param, max_chunk_cells, prune_log2_eps, clamp = ctx.cfg
        dcost = dcost.to(torch.float32).contiguous()
        d_enc, d_dec, dW, db = ops.joint_bwd(enc16, dec16, W16, b32, targets, ll, tl, blank, act_kind, act_param,
                                             lat2, logz, alpha, beta, cost, dcost, max_chunk_cells, prune_log2_eps, clamp)
        de, dd, dw, dbt = ctx.in_dtypes
        if d_enc.shape[-1] != ctx.H:  # drop the gradients of the zero padding
            d_enc, d_dec, dW = d_enc[..., :ctx.H].contiguous(), d_dec[..., :ctx.H].contiguous(), dW[:, :ctx.H].contiguous()
        return (d_enc.to(de), d_dec.to(dd), dW.to(dw), db.to(dbt)) + (None,) * 11


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
    if targets.device != enc_out.device:
        targets = targets.to(enc_out.device)
    costs, logit_lengths, _ = FusedJointRnnt.apply(enc_out, dec_out, weight, bias, targets, logit_lengths, target_lengths,
                                                bool(relative_lengths), int(blank), _lib.ACT_CODES[activation], float(act_param),
                                                int(max_chunk_cells), prune_log2_eps, float(clamp), bool(check_lengths))
    if numba_semantics:
        return _NumbaReduce.apply(costs, logit_lengths, reduction)
    return _reduce(costs, reduction)
