"""float64 numpy restatement of the reference's transducer loss path (TEST ORACLE ONLY).

Every function cites the reference lines it restates.  Paths are relative to /root/reference,
``SB/`` = ``vendor/speechbrain/speechbrain/``.

Two semantics exist in the reference:

* torchaudio semantics -- what the recipe runs: ``SB/nnet/losses.py:72-79`` calls
  ``torchaudio.functional.rnnt_loss`` (third-party, not vendored; ``torchaudio>=0.9.0`` in
  ``vendor/speechbrain/requirements.txt``, installed build 2.11.0+cu128).  Its published
  algorithm (Graves 2012, "Sequence Transduction with RNNs", eq. 16-20, with the log-softmax
  folded into the gradient as in warp-transducer) is restated in ``rnnt_torchaudio``.
* Numba semantics -- ``SB/nnet/loss/transducer_loss.py``: loss divided by T_b, gradient taken
  w.r.t. log-probs and left un-normalised; restated in ``rnnt_numba``.
"""
import numpy as np

NEG_INF = -np.inf


def lengths_from_relative(rel, dim):
    """``(rel * dim).round().int()`` -- SB/nnet/losses.py:58-59 (fp32 multiply, round-half-even)."""
    rel = np.asarray(rel, dtype=np.float32)
    return np.rint(rel * np.float32(dim)).astype(np.int32)


def log_softmax(x):
    """``logits.log_softmax(-1)`` -- SB/nnet/losses.py:84, SB/nnet/loss/transducer_loss.py:344."""
    x = np.asarray(x, dtype=np.float64)
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def _logaddexp(a, b):
    # max + log1p(exp(-|a-b|)) -- SB/nnet/loss/transducer_loss.py:94-96 and :171-173
    return np.logaddexp(a, b)


def two_value_lattice(lp, targets, blank):
    """skip(t,u) = lp[b,t,u,blank]; emit(t,u) = lp[b,t,u,targets[b,u]] for u < U-1.

    Indexing follows SB/nnet/loss/transducer_loss.py:81-90 (``labels[b, u-1]`` used when moving
    from column u-1 to u) and :160-166 (``labels[b, u]`` when moving from u to u+1).
    """
    B, T, U, V = lp.shape
    skip = lp[..., blank]
    emit = np.full((B, T, U), NEG_INF)
    if U > 1:
        idx = np.asarray(targets, dtype=np.int64)[:, None, :, None]  # [B,1,U-1,1]
        emit[:, :, : U - 1] = np.take_along_axis(lp[:, :, : U - 1, :], np.broadcast_to(idx, (B, T, U - 1, 1)), axis=-1)[..., 0]
    return skip, emit


def alpha_beta_one(skip, emit, Tb, Ub):
    """Forward/backward variables of one utterance on its Tb x Ub rectangle (Ub = labels+1).

    alpha: SB/nnet/loss/transducer_loss.py:60-99  (alpha[0,0]=0; first column adds blanks,
           first row adds labels, interior = logaddexp(no_emit, emit)).
    beta : SB/nnet/loss/transducer_loss.py:139-175 (beta[T-1,U]=lp_blank; mirror recursion).
    Returns (alpha, beta, logp) with logp = beta[0,0] = alpha[T-1,U-1] + skip[T-1,U-1].
    """
    alpha = np.full((Tb, Ub), NEG_INF)
    beta = np.full((Tb, Ub), NEG_INF)
    alpha[0, 0] = 0.0
    for t in range(Tb):
        for u in range(Ub):
            if t == 0 and u == 0:
                continue
            a = alpha[t - 1, u] + skip[t - 1, u] if t > 0 else NEG_INF
            b = alpha[t, u - 1] + emit[t, u - 1] if u > 0 else NEG_INF
            alpha[t, u] = _logaddexp(a, b)
    beta[Tb - 1, Ub - 1] = skip[Tb - 1, Ub - 1]
    for t in range(Tb - 1, -1, -1):
        for u in range(Ub - 1, -1, -1):
            if t == Tb - 1 and u == Ub - 1:
                continue
            a = beta[t + 1, u] + skip[t, u] if t < Tb - 1 else NEG_INF
            b = beta[t, u + 1] + emit[t, u] if u < Ub - 1 else NEG_INF
            beta[t, u] = _logaddexp(a, b)
    return alpha, beta, beta[0, 0]


def rnnt_torchaudio(logits, targets, logit_lengths, target_lengths, blank=0, clamp=-1.0):
    """torchaudio.functional.rnnt_loss(..., reduction="none", fused_log_softmax=True) semantics.

    Call site restated: SB/nnet/losses.py:72-79.  Returns (costs[B], dlogits[B,T,U,V]) with
    costs_b = -log P(y_b | x_b) and dlogits = d costs_b / d logits (softmax folded in):
        g[t,u,v] = p_v * exp(alpha+beta-L) - [v=blank] exp(alpha + lp_blank + beta(t+1,u) - L)
                                           - [v=y_u+1] exp(alpha + lp_v + beta(t,u+1) - L)
    with the terminal blank at (Tb-1,Ub-1) using beta := 0, and exactly 0 outside Tb x Ub.
    """
    logits = np.asarray(logits, dtype=np.float64)
    B, T, U, V = logits.shape
    if blank < 0:
        blank += V
    lp = log_softmax(logits)
    skip, emit = two_value_lattice(lp, targets, blank)
    costs = np.zeros(B)
    grads = np.zeros_like(logits)
    for b in range(B):
        Tb, Ub = int(logit_lengths[b]), int(target_lengths[b]) + 1
        alpha, beta, L = alpha_beta_one(skip[b], emit[b], Tb, Ub)
        costs[b] = -L
        p = np.exp(lp[b, :Tb, :Ub])  # [Tb,Ub,V]
        occ = np.exp(alpha + beta - L)  # [Tb,Ub]
        g = p * occ[..., None]
        beta_t1 = np.full((Tb, Ub), NEG_INF)
        beta_t1[:-1] = beta[1:]
        beta_t1[Tb - 1, Ub - 1] = 0.0
        g[..., blank] -= np.exp(alpha + skip[b, :Tb, :Ub] + beta_t1 - L)
        if Ub > 1:
            oe = np.exp(alpha[:, : Ub - 1] + emit[b, :Tb, : Ub - 1] + beta[:, 1:] - L)
            tu = np.asarray(targets[b, : Ub - 1], dtype=np.int64)
            for u in range(Ub - 1):
                g[:, u, tu[u]] -= oe[:, u]
        if clamp > 0:
            g = np.clip(g, -clamp, clamp)
        grads[b, :Tb, :Ub] = g
    return costs, grads


def reduce_costs(costs, reduction):
    """torchaudio functional.py:1791-1794 -- reduction applied outside the autograd Function."""
    if reduction == "mean":
        return costs.mean()
    if reduction == "sum":
        return costs.sum()
    if reduction == "none":
        return costs
    raise ValueError('reduction should be one of "none", "mean", or "sum"')


def rnnt_numba(log_probs, labels, T, U, blank=0, reduction="mean"):
    """``Transducer.apply(log_probs, labels, T, U, blank, reduction)`` semantics.

    SB/nnet/loss/transducer_loss.py:252-287: per-utterance value is -log P / T_b (:104-106),
    the reduction is applied inside forward (:280-287), and the stored gradient (w.r.t. the
    log-probs, :183-236) is -occupation at ``blank`` and at ``labels[b,u]``, zero elsewhere,
    NOT divided by T_b or B.  ``U`` here is the label count (lattice width - 1).
    Returns (loss, grads[B,T,U+1,V]).
    """
    lp = np.asarray(log_probs, dtype=np.float64)
    B, maxT, maxU, V = lp.shape
    skip, emit = two_value_lattice(lp, labels, blank)
    per_utt = np.zeros(B)
    grads = np.zeros_like(lp)
    for b in range(B):
        Tb, Ub = int(T[b]), int(U[b]) + 1
        alpha, beta, L = alpha_beta_one(skip[b], emit[b], Tb, Ub)
        per_utt[b] = -L / Tb
        beta_t1 = np.full((Tb, Ub), NEG_INF)
        beta_t1[:-1] = beta[1:]
        beta_t1[Tb - 1, Ub - 1] = 0.0
        grads[b, :Tb, :Ub, blank] = -np.exp(alpha + skip[b, :Tb, :Ub] + beta_t1 - L)
        for u in range(Ub - 1):
            l = int(labels[b, u])
            grads[b, :Tb, u, l] = -np.exp(alpha[:, u] + emit[b, :Tb, u] + beta[:, u + 1] - L)
    if reduction == "mean":
        loss = per_utt.mean()
    elif reduction == "sum":
        loss = per_utt.sum()
    elif reduction == "none":
        loss = per_utt
    else:
        raise Exception("Unexpected reduction {}".format(reduction))
    return loss, grads


# ----------------------------------------------------------------------------------------------
# Joint chain (Transducer_joint "sum" + nonlinearity -> Linear) and its backward, in float64.
# ----------------------------------------------------------------------------------------------

def bf16_round(x):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what the fused kernels feed the tensor core)."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    nan = np.isnan(x)
    out[nan] = np.nan
    return out


def activation(x, kind, param=0.01):
    """nonlinearity(joint) -- SB/nnet/transducer/transducer_joint.py:95 (LeakyReLU default :41)."""
    if kind == "leaky_relu":
        return np.where(x >= 0, x, x * param)
    if kind == "relu":
        return np.maximum(x, 0)
    if kind == "tanh":
        return np.tanh(x)
    if kind == "identity":
        return x
    raise ValueError(kind)


def activation_grad_from_output(j, kind, param=0.01):
    """act'(pre) expressed through the activation output j (sign-preserving activations)."""
    if kind == "leaky_relu":
        return np.where(j > 0, 1.0, param)
    if kind == "relu":
        return (j > 0).astype(j.dtype)
    if kind == "tanh":
        return 1.0 - j * j
    if kind == "identity":
        return np.ones_like(j)
    raise ValueError(kind)


def joint_logits(enc, dec, W, bias, act="leaky_relu", act_param=0.01, round_bf16=True):
    """joiner(enc[:,:,None,:], dec[:,None,:,:]) -> transducer_head  (fp64 accumulate).

    train_librispeechmix_scratch.py:132,135; SB/nnet/transducer/transducer_joint.py:73-74,95;
    SB/nnet/linear.py:74.  With ``round_bf16`` the operands are rounded exactly as the fused
    kernels round them: enc, dec, W to bf16 on entry; J = bf16(act(fp32(enc)+fp32(dec))).
    Returns (J[B,T,U,H] fp64, logits[B,T,U,V] fp64).
    """
    if round_bf16:
        enc, dec, W = bf16_round(enc), bf16_round(dec), bf16_round(W)
    pre = enc.astype(np.float32)[:, :, None, :] + dec.astype(np.float32)[:, None, :, :]
    J = activation(pre, act, np.float32(act_param)).astype(np.float32)
    if round_bf16:
        J = bf16_round(J)
    J = J.astype(np.float64)
    logits = J @ np.asarray(W, dtype=np.float64).T + np.asarray(bias, dtype=np.float64)
    return J, logits


def joint_backward(J, dlogits, W, act="leaky_relu", act_param=0.01, round_bf16=True):
    """Backward of Linear + activation + broadcast add (autograd of the chain above), fp64."""
    W = bf16_round(W).astype(np.float64) if round_bf16 else np.asarray(W, dtype=np.float64)
    dJ = dlogits @ W  # [B,T,U,H]
    dW = np.einsum("btuv,btuh->vh", dlogits, J)
    db = dlogits.sum(axis=(0, 1, 2))
    dpre = dJ * activation_grad_from_output(J, act, act_param)
    return dpre.sum(axis=2), dpre.sum(axis=1), dW, db
