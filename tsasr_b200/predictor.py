"""Drop-ins for the prediction network of the transducer.

* ``Embedding`` -- vendor/speechbrain/speechbrain/nnet/embedding.py:17-114 (same constructor, same frozen eye-matrix
  ``Embedding.weight`` so checkpoints load unchanged).  With ``consider_as_one_hot=True`` on CUDA token ids it returns a
  deferred ``OneHotHandle`` (virtual shape [B,U,V-1]) instead of materialising the one-hot tensor.
* ``LSTM`` -- vendor/speechbrain/speechbrain/nnet/RNN.py:170-278 (same constructor, the ``torch.nn.LSTM`` lives in
  ``self.rnn`` so ``rnn.weight_ih_l0`` ... keep their names; the class is NAMED ``LSTM``: the searchers recognise recurrent
  layers by class name, SB/decoders/transducer.py:491-499).  The teacher-forced call of the recipe --
  ``decoder(embedding(tokens_bos), lengths=tokens_bos_lens)``, train_librispeechmix_scratch.py:125-126 -- runs as ONE
  cooperative kernel (``tsasr_lstm_fwd``): the one-hot product with ``W_ih`` is a column gather, the relative lengths are
  converted on the device (the reference copies them to the host: ``.cpu()`` at SB/nnet/RNN.py:35, a stream
  synchronisation per step), padded positions come out as zeros exactly like ``pad_packed_sequence``.  Backward:
  ``tsasr_lstm_bwd`` (BPTT in one launch) + ``tsasr_linear_bwd`` (dW_hh, biases) + ``tsasr_onehot_dw`` (dW_ih).

Everything the kernels do not cover (initial state given -- the decode-time single steps --, several layers,
bidirectional, dropout between layers, hidden sizes other than 128/256/512, more than 64 utterances, CPU tensors) takes
the reference's own path through ``self.rnn`` (cuDNN), with the handle materialised by the reference's embedding lookup.

One deliberate difference: the reference's output is as long as the LONGEST utterance of the batch (``pad_packed_sequence``
needs the lengths on the host for that); the fused path always returns the padded length U.  SpeechBrain's relative
lengths make the longest utterance exactly 1.0, so the two coincide whenever the input is a PaddedBatch.
"""
import logging

import torch
import torch.nn as nn

from . import _lib
from .linear import _workspace as _linear_workspace
from .ops import _on_device

logger = logging.getLogger(__name__)

FUSED_HIDDEN_SIZES = (128, 256, 512)  # check_lstm_dims (csrc/predictor_capi.inl)
FUSED_MAX_BATCH = 64
FUSED_MAX_POSITIONS = 40000  # B*U W_ih column indices live in shared memory beside the 32 KB state tile (lstm_fwd_smem_bytes)


class OneHotHandle(torch.Tensor):
    """Deferred one-hot embedding of a [B,U] token tensor; never holds the [B,U,V-1] data."""

    @staticmethod
    def __new__(cls, tokens, weight, blank_id):
        r = torch.Tensor._make_wrapper_subclass(cls, (*tokens.shape, weight.shape[1]), dtype=weight.dtype, device=tokens.device,
                                                requires_grad=False)
        r._tokens, r._weight, r._blank_id = tokens, weight, blank_id
        return r

    def __init__(self, *args, **kwargs):
        pass

    def __repr__(self):
        return f"OneHotHandle(shape={tuple(self.shape)}, device={self.device})"

    def materialize(self):
        """Reference lookup (SB/nnet/embedding.py:114 on the frozen eye matrix, padding_idx = blank)."""
        return torch.nn.functional.embedding(self._tokens.long(), self._weight, padding_idx=self._blank_id)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name == "__get__" or func in _METADATA_FUNCS:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*_unwrap(tuple(args)), **{k: _unwrap(v) for k, v in kwargs.items()})

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        return func(*_unwrap(tuple(args)), **{k: _unwrap(v) for k, v in kwargs.items()})


def _unwrap(a):
    if isinstance(a, OneHotHandle):
        return a.materialize()
    if isinstance(a, (list, tuple)):
        return type(a)(_unwrap(v) for v in a)
    return a


_METADATA_FUNCS = {
    torch.Tensor.dim, torch.Tensor.size, torch.Tensor.ndimension, torch.Tensor.numel, torch.Tensor.is_floating_point,
    torch.Tensor.element_size, torch.Tensor.get_device, torch.Tensor.__len__, torch.Tensor.__repr__, torch.Tensor.__hash__,
    torch.Tensor.nelement,
}


class Embedding(nn.Module):
    """Computes an embedding x = wx.  Same arguments as the reference class (SB/nnet/embedding.py:65-103):
    num_embeddings, embedding_dim=128, consider_as_one_hot=False, blank_id=0."""

    def __init__(self, num_embeddings, embedding_dim=128, consider_as_one_hot=False, blank_id=0):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.consider_as_one_hot = consider_as_one_hot
        self.embedding_dim = self.num_embeddings - 1 if self.consider_as_one_hot else embedding_dim
        self.blank_id = blank_id
        if self.consider_as_one_hot:
            # the blank maps to the zero vector (padding_idx), every other token k to e_{k - [k > blank]}
            self.Embedding = nn.Embedding(self.num_embeddings, self.embedding_dim, padding_idx=self.blank_id)
            eye = torch.eye(self.embedding_dim)
            with torch.no_grad():
                self.Embedding.weight[self.blank_id + 1:] = eye[self.blank_id:]
                self.Embedding.weight[: self.blank_id] = eye[: self.blank_id]
                self.Embedding.weight[self.blank_id].zero_()
            self.Embedding.weight.requires_grad = False
        else:
            self.Embedding = nn.Embedding(self.num_embeddings, self.embedding_dim)

    def forward(self, x):
        """Returns the embedding of input tensor (a deferred handle for one-hot embeddings of CUDA token matrices)."""
        if self.consider_as_one_hot and x.is_cuda and x.dim() == 2 and not x.dtype.is_floating_point and x.dtype != torch.bool:
            return OneHotHandle(x, self.Embedding.weight, self.blank_id)
        return self.Embedding(x.long())  # pytorch embedding layer only accept long dtype


def _ptr(t):
    return t.data_ptr() if t is not None else None


class LstmFunction(torch.autograd.Function):
    """(out [B,U,Hd], h_n [B,Hd], c_n [B,Hd]) of a one-layer LSTM over a one-hot (tokens) or dense (x) input."""

    @staticmethod
    def forward(ctx, x, W_ih, W_hh, b_ih, b_hh, tokens, blank, lengths, relative):
        dev = W_hh.device
        lib = _lib.load()
        Hd = W_hh.shape[1]
        G = 4 * Hd
        onehot = tokens is not None
        B, U = (tokens.shape if onehot else x.shape[:2])
        stream = torch.cuda.current_stream(dev).cuda_stream
        need_grad = any(ctx.needs_input_grad[:5])
        out = torch.empty((B, U, Hd), dtype=torch.float32, device=dev)
        hprev = torch.empty((B, U, Hd), dtype=torch.float32, device=dev) if need_grad else None
        gates = torch.empty((B, U, G), dtype=torch.float32, device=dev) if need_grad else None
        cells = torch.empty((B, U, Hd), dtype=torch.float32, device=dev) if need_grad else None
        h_n = torch.empty((B, Hd), dtype=torch.float32, device=dev)
        c_n = torch.empty((B, Hd), dtype=torch.float32, device=dev)
        L = torch.empty((B,), dtype=torch.int32, device=dev)
        x2d = xw = None
        with _on_device(dev):
            if not onehot:  # dense input: x W_ih^T + b_ih through the projection GEMM, then the recurrence
                x2d = x.reshape(B * U, -1)
                if not x2d.is_contiguous():
                    x2d = x2d.contiguous()
                xw = torch.empty((B * U, G), dtype=torch.float32, device=dev)
                _lib.check(lib.tsasr_linear_fwd(x2d.data_ptr(), W_ih.data_ptr(), _ptr(b_ih), B * U, x2d.shape[1], G, xw.data_ptr(), None, stream))
            _lib.check(lib.tsasr_lstm_fwd(
                _ptr(tokens), 1 if onehot and tokens.dtype == torch.int64 else 0, int(blank), W_ih.shape[1], _ptr(xw), W_ih.data_ptr(),
                W_hh.data_ptr(), _ptr(b_ih), _ptr(b_hh), lengths.data_ptr() if relative else None, None if relative else lengths.data_ptr(),
                B, U, Hd, out.data_ptr(), _ptr(hprev), _ptr(gates), _ptr(cells), h_n.data_ptr(), c_n.data_ptr(), L.data_ptr(), stream))
        if need_grad:
            ctx.save_for_backward(x2d, W_ih, W_hh, tokens, hprev, gates, cells, L)
            ctx.cfg = (onehot, int(blank), b_ih is not None, b_hh is not None, tuple(x.shape) if x is not None else None)
        return out, h_n, c_n

    @staticmethod
    def backward(ctx, d_out, d_hn, d_cn):
        x2d, W_ih, W_hh, tokens, hprev, gates, cells, L = ctx.saved_tensors
        onehot, blank, has_bih, has_bhh, x_shape = ctx.cfg
        dev = W_hh.device
        lib = _lib.load()
        B, U, Hd = hprev.shape
        G = 4 * Hd
        stream = torch.cuda.current_stream(dev).cuda_stream

        def f32(t):
            return None if t is None else t.to(torch.float32).contiguous()

        d_out, d_hn, d_cn = f32(d_out), f32(d_hn), f32(d_cn)
        if d_out is None:
            d_out = torch.zeros((B, U, Hd), dtype=torch.float32, device=dev)
        dG = torch.empty((B, U, G), dtype=torch.float32, device=dev)
        dW_hh = torch.empty_like(W_hh)
        db = torch.empty((G,), dtype=torch.float32, device=dev)
        dW_ih = torch.empty_like(W_ih) if ctx.needs_input_grad[1] else None
        dx = None
        with _on_device(dev):
            _lib.check(lib.tsasr_lstm_bwd(d_out.data_ptr(), _ptr(d_hn), _ptr(d_cn), W_hh.data_ptr(), gates.data_ptr(), cells.data_ptr(),
                                          L.data_ptr(), B, U, Hd, dG.data_ptr(), stream))
            # dW_hh = dG^T h_prev, db_ih = db_hh = column sums of dG: one split-K GEMM with the row of ones
            ws2 = _linear_workspace(dev, B * U, Hd, G)
            _lib.check(lib.tsasr_linear_bwd(dG.data_ptr(), hprev.data_ptr(), None, B * U, Hd, G, None, dW_hh.data_ptr(), db.data_ptr(),
                                            ws2.data_ptr(), ws2.numel(), stream))
            if onehot:
                if dW_ih is not None:
                    _lib.check(lib.tsasr_onehot_dw(tokens.data_ptr(), 1 if tokens.dtype == torch.int64 else 0, blank, W_ih.shape[1],
                                                   dG.data_ptr(), B * U, G, dW_ih.data_ptr(), stream))
            else:
                In = x2d.shape[1]
                need_dx = ctx.needs_input_grad[0]
                if need_dx:
                    dx = torch.empty_like(x2d)
                if need_dx or dW_ih is not None:
                    if dW_ih is None:
                        dW_ih = torch.empty_like(W_ih)
                    ws3 = _linear_workspace(dev, B * U, In, G)
                    _lib.check(lib.tsasr_linear_bwd(dG.data_ptr(), x2d.data_ptr(), W_ih.data_ptr(), B * U, In, G, _ptr(dx), dW_ih.data_ptr(),
                                                    None, ws3.data_ptr(), ws3.numel(), stream))
                if dx is not None:
                    dx = dx.view(x_shape)
        return (dx, dW_ih if ctx.needs_input_grad[1] else None, dW_hh, db if has_bih else None, db.clone() if has_bhh else None,
                None, None, None, None)


_warned = set()


def _warn_once(msg):
    if msg not in _warned:
        _warned.add(msg)
        logger.warning("tsasr_b200: %s", msg)


class LSTM(nn.Module):
    """This function implements a basic LSTM.  Input tensors are formatted as (batch, time, fea); 4-D inputs
    (batch, time, fea, channel) are flattened to (batch, time, fea*channel).

    Same arguments as the reference class (SB/nnet/RNN.py:213-252): hidden_size, input_shape / input_size, num_layers,
    bias, dropout, re_init (orthogonal recurrent weights), bidirectional."""

    def __init__(self, hidden_size, input_shape=None, input_size=None, num_layers=1, bias=True, dropout=0.0, re_init=True,
                 bidirectional=False):
        super().__init__()
        self.reshape = False
        if input_shape is None and input_size is None:
            raise ValueError("Expected one of input_shape or input_size.")
        if input_size is None:  # computing the feature dimensionality
            if len(input_shape) > 3:
                self.reshape = True
            input_size = torch.prod(torch.tensor(input_shape[2:])).item()
        self.rnn = torch.nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, dropout=dropout,
                                 bidirectional=bidirectional, bias=bias, batch_first=True)
        if re_init:  # rnn_init, SB/nnet/RNN.py:2024-2042: orthogonal recurrent weights
            for name, param in self.rnn.named_parameters():
                if "weight_hh" in name or ".u.weight" in name:
                    nn.init.orthogonal_(param)

    def _fused_ok(self, x, hx):
        r = self.rnn
        if hx is not None or r.num_layers != 1 or r.bidirectional or r.proj_size != 0:
            return False
        if r.hidden_size not in FUSED_HIDDEN_SIZES:
            _warn_once(f"LSTM hidden size {r.hidden_size} is outside {FUSED_HIDDEN_SIZES}: using the reference's cuDNN path")
            return False
        if x.dim() != 3 or x.shape[1] < 2 or x.shape[0] < 1:
            return False  # single steps (decoding) and degenerate shapes: the module's own torch.nn.LSTM
        if x.shape[0] > FUSED_MAX_BATCH:
            _warn_once(f"{x.shape[0]} utterances exceed the {FUSED_MAX_BATCH} the fused recurrence keeps state for: using the "
                       "reference's cuDNN path")
            return False
        if isinstance(x, OneHotHandle) and x.shape[0] * x.shape[1] > FUSED_MAX_POSITIONS:
            _warn_once(f"{x.shape[0]} x {x.shape[1]} token positions exceed the shared-memory table of the fused recurrence "
                       f"({FUSED_MAX_POSITIONS}): using the reference's cuDNN path")
            return False
        w = r.weight_hh_l0
        if not (x.is_cuda and w.is_cuda and w.dtype == torch.float32) or torch.is_autocast_enabled():
            return False
        return isinstance(x, OneHotHandle) or x.dtype == torch.float32

    def forward(self, x, hx=None, lengths=None):
        """Returns the output of the LSTM: (output [B,U,hidden], (h_n, c_n)).  ``lengths``: RELATIVE lengths."""
        if self.reshape and x.ndim == 4:  # reshaping input tensors for 4d inputs
            x = x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3])
        if self._fused_ok(x, hx):
            r = self.rnn
            dev = r.weight_hh_l0.device
            B, U = x.shape[0], x.shape[1]
            if lengths is None:
                lens, relative = torch.full((B,), U, dtype=torch.int32, device=dev), False
            else:
                lens, relative = lengths.to(device=dev, dtype=torch.float32).contiguous(), True
            b_ih = r.bias_ih_l0 if r.bias else None
            b_hh = r.bias_hh_l0 if r.bias else None
            if isinstance(x, OneHotHandle):
                tokens = x._tokens if x._tokens.dtype in (torch.int32, torch.int64) else x._tokens.long()
                out, h_n, c_n = LstmFunction.apply(None, r.weight_ih_l0, r.weight_hh_l0, b_ih, b_hh, tokens.contiguous(), x._blank_id,
                                                   lens, relative)
            else:
                out, h_n, c_n = LstmFunction.apply(x, r.weight_ih_l0, r.weight_hh_l0, b_ih, b_hh, None, 0, lens, relative)
            return out, (h_n.unsqueeze(0), c_n.unsqueeze(0))

        # ---- the reference's own path (SB/nnet/RNN.py:254-278) ----
        if isinstance(x, OneHotHandle):
            x = x.materialize()
        self.rnn.flatten_parameters()  # flatten params for data parallel
        if lengths is not None:  # pack sequence for proper RNN handling of padding
            x = torch.nn.utils.rnn.pack_padded_sequence(x, (lengths * x.size(1)).cpu(), batch_first=True, enforce_sorted=False)
        output, hn = self.rnn(x, hx=hx) if hx is not None else self.rnn(x)
        if lengths is not None:  # unpack the packed sequence
            output, _ = torch.nn.utils.rnn.pad_packed_sequence(output, batch_first=True)
        return output, hn
