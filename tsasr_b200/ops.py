"""Thin torch-tensor wrappers over the C ABI (include/tsasr_b200.h).

torch is used for device memory, streams and dtype bookkeeping only; all arithmetic happens in
libtsasr_b200.so.  Every wrapper launches on ``torch.cuda.current_stream()`` of the tensors' device
and never synchronises.
"""
import ctypes

import torch

from . import _lib

_DTYPE_CODE = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _p(t):
    # plain integers: the argtypes in _lib.SIGNATURES convert them (a ctypes object per pointer costs ~0.5 us, and the
    # host time in front of the first kernel of a step is GPU idle time)
    return t.data_ptr() if t is not None else None


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """``with torch.cuda.device(dev)`` only when ``dev`` is not already current (the context manager costs ~10 us of
    host time per call, which is exposed in front of the first kernel of a step)."""

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def _require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError(f"tsasr_b200 needs CUDA tensors (got a tensor on {t.device}); there is no CPU path")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors live on different devices: {dev} vs {t.device}")
    return dev


def lattice_elems(B, T, U):
    return B * (T + U - 1) * U


def logits_to_lattice(logits, targets, logit_lengths, target_lengths, blank, normalized=False):
    """[B,T,U,V] logits (or log-probs) -> (lat2 [cells,2] fp32, den [cells] fp32), skewed layout."""
    dev = _require_cuda(logits, targets, logit_lengths, target_lengths)
    B, T, U, V = logits.shape
    n = lattice_elems(B, T, U)
    lat2 = torch.empty((n, 2), dtype=torch.float32, device=dev)
    den = torch.empty((n,), dtype=torch.float32, device=dev)
    with _on_device(dev):
        _lib.check(_lib.load().tsasr_logits_to_lattice(
            _p(logits), _DTYPE_CODE[logits.dtype], _p(targets), _p(logit_lengths), _p(target_lengths),
            B, T, U, V, int(blank), int(bool(normalized)), _p(lat2), _p(den), _stream(dev)))
    return lat2, den


def alpha_beta(lat2, logit_lengths, target_lengths, B, T, U):
    """Wavefront DP -> (alpha, beta [cells] fp32, cost [B] = -log P, ll_alpha [B], ll_beta [B])."""
    dev = _require_cuda(lat2, logit_lengths, target_lengths)
    n = lattice_elems(B, T, U)
    buf = torch.empty((2 * n + 3 * B,), dtype=torch.float32, device=dev)  # one allocation: alpha | beta | cost, ll_a, ll_b
    alpha, beta, out = buf[:n], buf[n: 2 * n], buf[2 * n:].view(3, B)
    with _on_device(dev):
        _lib.check(_lib.load().tsasr_lattice_alpha_beta(
            _p(lat2), _p(logit_lengths), _p(target_lengths), B, T, U, _p(alpha), _p(beta),
            _p(out[0]), _p(out[1]), _p(out[2]), _stream(dev)))
    return alpha, beta, out[0], out[1], out[2]


def logits_grad(logits, targets, logit_lengths, target_lengths, blank, lat2, den, alpha, beta, cost, dcost, clamp=-1.0):
    dev = _require_cuda(logits, lat2, den, alpha, beta, cost, dcost)
    B, T, U, V = logits.shape
    dlogits = torch.empty_like(logits)
    with _on_device(dev):
        _lib.check(_lib.load().tsasr_logits_grad(
            _p(logits), _DTYPE_CODE[logits.dtype], _p(targets), _p(logit_lengths), _p(target_lengths),
            B, T, U, V, int(blank), _p(lat2), _p(den), _p(alpha), _p(beta), _p(cost), _p(dcost),
            float(clamp), _p(dlogits), _stream(dev)))
    return dlogits


def logprobs_grad(shape, targets, logit_lengths, target_lengths, blank, lat2, alpha, beta, cost, dcost):
    dev = _require_cuda(lat2, alpha, beta, cost, dcost)
    B, T, U, V = shape
    grads = torch.empty(shape, dtype=torch.float32, device=dev)
    with _on_device(dev):
        _lib.check(_lib.load().tsasr_logprobs_grad(
            _p(targets), _p(logit_lengths), _p(target_lengths), B, T, U, V, int(blank), _p(lat2), _p(alpha),
            _p(beta), _p(cost), _p(dcost), _p(grads), _stream(dev)))
    return grads


def joint_fwd(enc, dec, W, bias, targets, logit_lengths, target_lengths, blank, act_kind, act_param):
    """bf16 enc [B,T,H], dec [B,U,H], W [V,H]; fp32 bias [V] -> (lat2, logz), skewed layout."""
    dev = _require_cuda(enc, dec, W, bias, targets, logit_lengths, target_lengths)
    B, T, H = enc.shape
    U = dec.shape[1]
    V = W.shape[0]
    n = lattice_elems(B, T, U)
    buf = torch.empty((3 * n,), dtype=torch.float32, device=dev)  # one allocation: lat2 | logz
    lat2, logz = buf[: 2 * n].view(n, 2), buf[2 * n:]
    with _on_device(dev):
        _lib.check(_lib.load().tsasr_joint_fwd(
            _p(enc), _p(dec), _p(W), _p(bias), _p(targets), _p(logit_lengths), _p(target_lengths),
            B, T, U, H, V, int(blank), int(act_kind), float(act_param), _p(lat2), _p(logz), _stream(dev)))
    return lat2, logz


_workspaces = {}          # (device, stream handle) -> uint8 workspace tensor, least recently used first
_MAX_CACHED_WORKSPACES = 4  # per process: a stream that stops running backward passes gives its buffer back


def _workspace(dev, nbytes):
    """Cached workspace per (device, stream) -- two backward passes queued on different streams of one device must not
    share their operand images -- grown on demand; the torch caching allocator owns the memory.  The cache is bounded
    (least recently used entries are dropped, their memory returns to the allocator once the queued kernels are done:
    the allocator is stream-aware) and a buffer that is outgrown is released BEFORE its replacement is allocated."""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.pop(key, None)
    if ws is None or ws.numel() < nbytes:
        last = _last_bwd.get(dev)
        if last is not None and last[0] is ws:  # the statistics accessor must not pin the outgrown buffer
            del _last_bwd[dev]
        ws = last = None
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    _workspaces[key] = ws  # most recently used last
    while len(_workspaces) > _MAX_CACHED_WORKSPACES:
        old_key = next(iter(_workspaces))
        old = _workspaces.pop(old_key)
        last = _last_bwd.get(old_key[0])
        if last is not None and last[0] is old:
            del _last_bwd[old_key[0]]
        del old, last
    return ws


def last_workspace(dev):
    """The workspace the last joint_bwd on ``dev`` used (test / debugging helper: its head holds the operand images)."""
    return _last_bwd[dev][0]


def default_prune_log2_eps():
    """Backward tile pruning threshold (log2 of the smallest alignment posterior a tile must reach to be kept);
    TSASR_PRUNE_LOG2_EPS overrides the default -30, any value >= 0 switches pruning off."""
    import os

    return float(os.environ.get("TSASR_PRUNE_LOG2_EPS", "-30"))


_last_bwd = {}


def last_backward_tile_stats(dev=None):
    """(active tiles, live tiles) of the last joint_bwd on ``dev`` (synchronises; measurement aid)."""
    if dev is None:
        dev = next(iter(_last_bwd))
    ws, off = _last_bwd[dev]
    base = (-ws.data_ptr()) % 1024 + off
    st = ws[base: base + 12].view(torch.int32).tolist()
    return st[1], st[2]


def joint_bwd(enc, dec, W, bias, targets, logit_lengths, target_lengths, blank, act_kind, act_param,
              lat2, logz, alpha, beta, cost, dcost, max_chunk_cells=0, prune_log2_eps=None, clamp=-1.0):
    """Backward of the fused chain -> (d_enc [B,T,H], d_dec [B,U,H], dW [V,H], db [V]) fp32.
    ``clamp`` > 0: torchaudio's gradient clamp on the dlogits of the unit cost (see include/tsasr_b200.h)."""
    dev = _require_cuda(enc, dec, W, bias, lat2, logz, alpha, beta, cost, dcost)
    B, T, H = enc.shape
    U = dec.shape[1]
    V = W.shape[0]
    lib = _lib.load()
    nbytes = lib.tsasr_joint_bwd_workspace_bytes(B, T, U, H, V, int(max_chunk_cells))
    ws = _workspace(dev, max(int(nbytes), 16))
    if prune_log2_eps is None:
        prune_log2_eps = default_prune_log2_eps()
    _last_bwd[dev] = (ws, int(lib.tsasr_joint_bwd_stats_offset(B, T, U, H, V, int(max_chunk_cells))))
    d_enc = torch.empty((B, T, H), dtype=torch.float32, device=dev)
    d_dec = torch.empty((B, U, H), dtype=torch.float32, device=dev)
    dW = torch.empty((V, H), dtype=torch.float32, device=dev)
    db = torch.empty((V,), dtype=torch.float32, device=dev)
    with _on_device(dev):
        _lib.check(lib.tsasr_joint_bwd(
            _p(enc), _p(dec), _p(W), _p(bias), _p(targets), _p(logit_lengths), _p(target_lengths),
            B, T, U, H, V, int(blank), int(act_kind), float(act_param), _p(lat2), _p(logz), _p(alpha), _p(beta),
            _p(cost), _p(dcost), _p(ws), ctypes.c_size_t(ws.numel()), int(max_chunk_cells), float(prune_log2_eps), float(clamp),
            _p(d_enc), _p(d_dec), _p(dW), _p(db), _stream(dev)))
    return d_enc, d_dec, dW, db


def joint_debug_logits(enc, dec, W, bias, act_kind, act_param):
    """Test-only: dense fp32 logits recomputed by the tcgen05 mainloop (small shapes)."""
    dev = _require_cuda(enc, dec, W, bias)
    B, T, H = enc.shape
    U = dec.shape[1]
    V = W.shape[0]
    out = torch.zeros((B, T, U, V), dtype=torch.float32, device=dev)
    with _on_device(dev):
        _lib.check(_lib.load().tsasr_joint_debug_logits(
            _p(enc), _p(dec), _p(W), _p(bias), B, T, U, H, V, int(act_kind), float(act_param), _p(out), _stream(dev)))
    return out


def unskew(x, B, T, U):
    """Skewed lattice array [B*(T+U-1)*U, ...] -> dense [B,T,U,...] (test / debugging helper)."""
    D = T + U - 1
    x = x.reshape(B, D, U, *x.shape[1:])
    t = torch.arange(T, device=x.device)[:, None]
    u = torch.arange(U, device=x.device)[None, :]
    return x[:, t + u, u]
