"""Drop-in ``Linear`` for the projections either side of the joint.

Mirrors vendor/speechbrain/speechbrain/nnet/linear.py:18-76 (same constructor, same ``forward``, the ``nn.Linear`` lives
in ``self.w`` so checkpoints keep their ``w.weight`` / ``w.bias`` names) as the recipe instantiates it for
``encoder_proj`` and ``decoder_proj`` (hparams/LibriSpeechMix/conformer-t_scratch.yaml:172-174,187-189; called at
train_librispeechmix_scratch.py:122,127).

On CUDA fp32 inputs the product runs in ``tsasr_linear_fwd`` (tcgen05 GEMM on in-kernel bf16 hi/lo splits: 16-17 bits
per term) whose epilogue writes the fp32 output AND its bf16 rounding.  The bf16 copy is remembered here, keyed by the
output's storage, and ``FusedJointRnnt`` picks it up as the joint GEMM's operand image: the fp32 -> bf16 pass over
``enc_out`` / ``dec_out`` inside the fused loss disappears, and the backward (``tsasr_linear_bwd``) consumes the fp32
``d_enc`` / ``d_dec`` of the joint backward as they are.  A ``JointHandle`` input (this class used as the transducer
head) is forwarded to ``F.linear`` so that the handle attaches ``W, b`` exactly as with the stock head.  CPU tensors and
non-fp32 dtypes use the reference's own op (``F.linear``), like every non-deferred case of ``Transducer_joint``.
"""
import collections

import torch
import torch.nn as nn

from . import _lib
from .ops import _on_device
from .transducer_joint import JointHandle

_TWINS_PER_DEVICE = 4
_twins = {}  # device -> OrderedDict{storage data_ptr: (detached fp32 output, its version counter, bf16 copy)}


def _remember_twin(y32, y16):
    reg = _twins.setdefault(y32.device, collections.OrderedDict())
    # the detached alias keeps the storage alive (so the address cannot be reused while the entry exists) without
    # keeping the autograd graph of the step alive; it shares the version counter with every view of the output
    alias = y32.detach()
    reg[alias.untyped_storage().data_ptr()] = (alias, alias._version, y16)
    while len(reg) > _TWINS_PER_DEVICE:
        reg.popitem(last=False)


def bf16_twin(t, consume=False):
    """The bf16 copy written together with ``t`` by ``tsasr_linear_fwd``, if ``t`` is (a full, contiguous view of) such
    an output and has not been modified in place since; else None.  ``consume``: drop the registry entry (the fused loss
    keeps the copy alive through its autograd node from then on, so nothing outlives the step)."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.storage_offset() == 0):
        return None
    reg = _twins.get(t.device)
    if not reg:
        return None
    key = t.untyped_storage().data_ptr()
    hit = reg.get(key)
    if hit is None:
        return None
    alias, version, y16 = hit
    if alias.numel() != t.numel() or alias._version != version:
        return None
    if consume:
        del reg[key]
    return y16.view(t.shape)


def forget_twins():
    """Drop the remembered bf16 copies (they keep at most ``_TWINS_PER_DEVICE`` projection outputs per device alive)."""
    _twins.clear()


_ws_bytes = {}   # (R, K, N) -> tsasr_linear_bwd_workspace_bytes
_ws_cache = {}   # (device, stream) -> split-K workspace, grown on demand (the host time of a small GEMM's backward is
                 # mostly allocator calls and context managers otherwise)


def _workspace(dev, R, K, N):
    key = (R, K, N)
    n = _ws_bytes.get(key)
    if n is None:
        n = _ws_bytes[key] = max(int(_lib.load().tsasr_linear_bwd_workspace_bytes(R, K, N)), 256)
    ck = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(ck)
    if ws is None or ws.numel() < n:
        if len(_ws_cache) >= 8:
            _ws_cache.clear()
        ws = _ws_cache[ck] = torch.empty((n,), dtype=torch.uint8, device=dev)
    return ws


class LinearFunction(torch.autograd.Function):
    """(Y fp32, Y bf16) = X W^T + b through tsasr_linear_fwd; backward through tsasr_linear_bwd."""

    @staticmethod
    def forward(ctx, x2d, weight, bias):
        R, K = x2d.shape
        N = weight.shape[0]
        dev = x2d.device
        lib = _lib.load()
        y = torch.empty((R, N), dtype=torch.float32, device=dev)
        y16 = torch.empty((R, N), dtype=torch.bfloat16, device=dev)
        with _on_device(dev):
            _lib.check(lib.tsasr_linear_fwd(x2d.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                            R, K, N, y.data_ptr(), y16.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(x2d, weight)
        ctx.has_bias = bias is not None
        ctx.mark_non_differentiable(y16)
        return y, y16

    @staticmethod
    def backward(ctx, dy, _dy16=None):
        x2d, weight = ctx.saved_tensors
        R, K = x2d.shape
        N = weight.shape[0]
        dev = x2d.device
        lib = _lib.load()
        if dy.dtype != torch.float32 or not dy.is_contiguous():
            dy = dy.to(torch.float32).contiguous()
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_db = ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x2d) if need_dx else None
        dw = torch.empty_like(weight) if (need_dw or need_db) else None
        db = torch.empty((N,), dtype=torch.float32, device=dev) if need_db else None
        if dx is None and dw is None:
            return None, None, None
        ws = _workspace(dev, R, K, N) if dw is not None else None
        with _on_device(dev):
            _lib.check(lib.tsasr_linear_bwd(dy.data_ptr(), x2d.data_ptr(), weight.data_ptr(), R, K, N,
                                            dx.data_ptr() if dx is not None else None, dw.data_ptr() if dw is not None else None,
                                            db.data_ptr() if db is not None else None, ws.data_ptr() if ws is not None else None,
                                            ws.numel() if ws is not None else 0, torch.cuda.current_stream(dev).cuda_stream))
        return dx, (dw if need_dw else None), db


def linear(x, weight, bias=None):
    """``F.linear(x, weight, bias)`` on the last dimension through the tcgen05 kernels (CUDA, fp32); the bf16 copy of the
    result is remembered for the fused joint (``bf16_twin``)."""
    if isinstance(x, JointHandle):
        return torch.nn.functional.linear(x, weight, bias)  # the handle attaches W, b (transducer head)
    fused = (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.is_cuda
             and (bias is None or bias.dtype == torch.float32) and x.numel() > 0 and not torch.is_autocast_enabled())
    if not fused:
        return torch.nn.functional.linear(x, weight, bias)
    K = x.shape[-1]
    x2d = x.reshape(-1, K)
    if not x2d.is_contiguous():
        x2d = x2d.contiguous()
    w = weight if weight.is_contiguous() else weight.contiguous()
    b = bias if bias is None or bias.is_contiguous() else bias.contiguous()
    y, y16 = LinearFunction.apply(x2d, w, b)
    y = y.view(*x.shape[:-1], weight.shape[0])
    _remember_twin(y, y16)
    return y


class Linear(nn.Module):
    """Computes a linear transformation y = wx + b.

    Same arguments as the reference class (SB/nnet/linear.py:18-61):
    n_neurons : output size;  input_shape / input_size : expected input (one of them is required);
    bias : add the bias term;  combine_dims : flatten the last two dimensions of a 4-D input first.
    """

    def __init__(self, n_neurons, input_shape=None, input_size=None, bias=True, combine_dims=False):
        super().__init__()
        self.combine_dims = combine_dims
        if input_shape is None and input_size is None:
            raise ValueError("Expected one of input_shape or input_size")
        if input_size is None:
            input_size = input_shape[-1]
            if len(input_shape) == 4 and self.combine_dims:
                input_size = input_shape[2] * input_shape[3]
        self.w = nn.Linear(input_size, n_neurons, bias=bias)  # weights are initialized following pytorch approach

    def forward(self, x):
        """Returns the linear transformation of input tensor (last dimension)."""
        if x.ndim == 4 and self.combine_dims:
            x = x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3])
        return linear(x, self.w.weight, self.w.bias)
