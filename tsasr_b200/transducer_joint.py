"""Drop-in ``Transducer_joint`` whose training-shape output is a deferred handle.

Mirrors vendor/speechbrain/speechbrain/nnet/transducer/transducer_joint.py:15-95 (same constructor,
same ``forward(input_TN, input_PN)``, same ``init_params``, no parameters and no buffers so
reference checkpoints load with strict=True).

For ``joint="sum"`` on CUDA inputs of shape [B,T,1,H] + [B,1,U,H] the module does NOT materialise
the [B,T,U,H] tensor: it returns a ``JointHandle`` -- a storage-less tensor subclass with the
virtual shape [B,T,U,H] that carries ``enc``, ``dec`` and the activation.  The stock
``speechbrain.nnet.linear.Linear`` head then calls ``torch.nn.functional.linear(handle, W, b)``
(SB/nnet/linear.py:74), which the handle intercepts to attach ``W, b`` (virtual shape [B,T,U,V]);
this happens inside the head's own ``DistributedDataParallel.forward``, so its reducer is armed
exactly as in the reference (SB/core.py:1469-1484).  ``transducer_loss`` consumes the handle and
launches the fused kernels.  Any other operation on a handle materialises it with the
reference's eager math (transducer_joint.py:74,95 and linear.py:74) -- that is what the greedy /
beam searchers hit on their tiny [B,1,1,H] inputs (SB/decoders/transducer.py:375-384).
"""
import logging

import torch
import torch.nn as nn

from . import _lib

logger = logging.getLogger(__name__)

_MIN_DEFERRED_CELLS = 2  # a single cell (decode-time [B,1,1,H] calls) uses the eager math
MAX_FUSED_H = 640        # kMaxKB * 64 (csrc/joint_gemm.cuh): the A operand of a cell tile is shared-memory resident
_warned = set()


def _warn_once(msg):
    """One warning per distinct message and process: a shape the fused path does not cover must not degrade silently."""
    if msg not in _warned:
        _warned.add(msg)
        logger.warning("tsasr_b200: %s", msg)


def activation_code(module):
    """(code, param) for the activations the fused prologue implements, else None."""
    if type(module) is nn.LeakyReLU:
        return _lib.ACT_CODES["leaky_relu"], float(module.negative_slope)
    if type(module) is nn.ReLU:
        return _lib.ACT_CODES["relu"], 0.0
    if type(module) is nn.Tanh:
        return _lib.ACT_CODES["tanh"], 0.0
    if type(module) is nn.Identity:
        return _lib.ACT_CODES["identity"], 0.0
    return None


class JointHandle(torch.Tensor):
    """Deferred ``act(enc + dec)`` (optionally followed by the head Linear); never holds B*T*U data."""

    @staticmethod
    def __new__(cls, enc, dec, act_module, act_code, act_param, weight=None, bias=None):
        B, T, _, H = enc.shape
        U = dec.shape[2]
        last = H if weight is None else weight.shape[0]
        r = torch.Tensor._make_wrapper_subclass(cls, (B, T, U, last), dtype=enc.dtype, device=enc.device,
                                                requires_grad=False)
        r._enc, r._dec = enc, dec
        r._act_module, r._act_code, r._act_param = act_module, act_code, act_param
        r._weight, r._bias = weight, bias
        return r

    def __init__(self, *args, **kwargs):
        pass

    def __repr__(self):
        return f"JointHandle(shape={tuple(self.shape)}, dtype={self.dtype}, device={self.device}, has_head={self._weight is not None})"

    @property
    def has_head(self):
        return self._weight is not None

    def materialize(self):
        """Reference eager math (transducer_joint.py:74,95; linear.py:74)."""
        joint = self._act_module(self._enc + self._dec)
        if self._weight is not None:
            joint = torch.nn.functional.linear(joint, self._weight, self._bias)
        return joint

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.nn.functional.linear:
            x = args[0] if args else kwargs.get("input")
            if isinstance(x, JointHandle) and not x.has_head:
                w = args[1] if len(args) > 1 else kwargs.get("weight")
                b = args[2] if len(args) > 2 else kwargs.get("bias")
                if (isinstance(w, torch.Tensor) and not isinstance(w, JointHandle) and w.dim() == 2 and w.shape[1] == x.shape[-1]
                        and w.shape[0] >= 2):  # V = 1 (blank only) has no fused kernel: materialise
                    return JointHandle(x._enc, x._dec, x._act_module, x._act_code, x._act_param, w, b)
        name = getattr(func, "__name__", "")
        if name == "__get__" or func in _METADATA_FUNCS:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        # anything else: materialise and run on plain tensors
        def unwrap(a):
            if isinstance(a, JointHandle):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(v) for v in a)
            return a

        with torch._C.DisableTorchFunctionSubclass():
            return func(*unwrap(tuple(args)), **{k: unwrap(v) for k, v in kwargs.items()})

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}

        def unwrap(a):
            if isinstance(a, JointHandle):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(v) for v in a)
            return a

        return func(*unwrap(tuple(args)), **{k: unwrap(v) for k, v in kwargs.items()})


_METADATA_FUNCS = {
    torch.Tensor.dim, torch.Tensor.size, torch.Tensor.ndimension, torch.Tensor.numel, torch.Tensor.stride,
    torch.Tensor.is_contiguous, torch.Tensor.is_floating_point, torch.Tensor.is_complex, torch.Tensor.element_size,
    torch.Tensor.get_device, torch.Tensor.__len__, torch.Tensor.__repr__, torch.Tensor.__hash__,
    torch.Tensor.data_ptr, torch.Tensor.storage_offset, torch.Tensor.nelement, torch.Tensor.is_shared,
    torch.Tensor.dim_order,
}


class Transducer_joint(nn.Module):
    """Computes joint tensor between Transcription network (TN) & Prediction network (PN).

    Same arguments as the reference class (transducer_joint.py:40-46):
    joint_network : module called after the concatenation when joint == "concat" (or None)
    joint : "sum" or "concat"
    nonlinearity : activation *class* (instantiated here), default torch.nn.LeakyReLU
    """

    def __init__(self, joint_network=None, joint="sum", nonlinearity=torch.nn.LeakyReLU):
        super().__init__()
        self.joint_network = joint_network
        self.joint = joint
        self.nonlinearity = nonlinearity()

    def init_params(self, first_input):
        self.joint_network(first_input)

    def _concat_as_sum(self, input_TN, input_PN):
        """``joint="concat"`` with a single Linear ``joint_network`` (transducer_joint.py:76-93) IS a "sum" joint of two
        projections: ``W_j [e; d] + b = (W_j[:, :He] e + b) + W_j[:, He:] d``.  The reference expands both tensors to
        [B,T,U,He+Hd], concatenates and runs the Linear over all B*T*U rows; here the two halves of ``W_j`` are applied to
        the B*T encoder rows and the B*U predictor rows (plain ``F.linear``: autograd carries the gradient back into
        ``joint_network``), and the rest is the fused "sum" path.  Returns the projected pair or None."""
        if self.joint != "concat" or self.joint_network is None or input_TN.dim() != 4:
            return None
        lin = getattr(self.joint_network, "w", self.joint_network)  # SpeechBrain's Linear keeps its nn.Linear in .w
        if type(lin) is not nn.Linear or not (input_TN.is_cuda and input_PN.is_cuda):
            return None
        if getattr(self.joint_network, "combine_dims", False):
            return None
        He, Hd = input_TN.shape[-1], input_PN.shape[-1]
        if lin.in_features != He + Hd or input_TN.shape[2] != 1 or input_PN.shape[1] != 1:
            return None
        W = lin.weight
        return torch.nn.functional.linear(input_TN, W[:, :He], lin.bias), torch.nn.functional.linear(input_PN, W[:, He:])

    def _deferrable(self, input_TN, input_PN, as_sum=False):
        if (self.joint != "sum" and not as_sum) or input_TN.dim() != 4 or not (input_TN.is_cuda and input_PN.is_cuda):
            return None
        if input_TN.shape[2] != 1 or input_PN.shape[1] != 1 or input_TN.shape[0] != input_PN.shape[0]:
            return None
        H = input_TN.shape[3]
        if input_PN.shape[3] != H or H < 1:
            return None
        if H > MAX_FUSED_H:  # H % 64 != 0 is zero-padded to whole k-blocks by the fused op
            _warn_once(f"joint dimension H = {H} > {MAX_FUSED_H}: the fused tcgen05 path keeps the whole joint operand of a cell "
                       "tile in shared memory and does not implement this width; falling back to the reference's eager math "
                       "(materialised [B,T,U,H] / [B,T,U,V] tensors, cuBLAS GEMMs, compat loss kernels)")
            return None
        if input_TN.shape[1] * input_PN.shape[2] < _MIN_DEFERRED_CELLS:
            return None
        if input_TN.dtype != input_PN.dtype or not input_TN.dtype.is_floating_point:
            return None
        return activation_code(self.nonlinearity)

    def forward(self, input_TN, input_PN):
        """Returns the fusion of inputs tensors (a deferred handle for training-shape CUDA inputs)."""
        if len(input_TN.shape) != len(input_PN.shape):
            raise ValueError("Arg 1 and 2 must be have same size")
        if not (len(input_TN.shape) != 4 or len(input_TN.shape) != 1):
            raise ValueError("Tensors 1 and 2 must have dim=1 or dim=4")  # tautology kept from the reference (:70)

        act = self._deferrable(input_TN, input_PN)
        if act is not None:
            _lib.load()  # fail loudly here if the CUDA extension is missing
            return JointHandle(input_TN, input_PN, self.nonlinearity, act[0], act[1])
        pair = self._concat_as_sum(input_TN, input_PN)
        if pair is not None:
            act = self._deferrable(*pair, as_sum=True)  # the deferral rules of the "sum" path apply to the projected pair
            if act is not None:
                _lib.load()
                return JointHandle(pair[0], pair[1], self.nonlinearity, act[0], act[1])

        return self._eager(input_TN, input_PN)

    def _eager(self, tn, pn):
        """The reference's eager math for every non-deferred case (decode-time shapes, CPU tensors, ``concat`` joints
        without a single-Linear ``joint_network``, activations the fused prologue does not implement)."""
        if self.joint == "sum":
            joint = tn + pn  # broadcast add, transducer_joint.py:73-74
        elif self.joint == "concat":
            if tn.dim() == 4:  # training: expand both to the common [B,T,U] lead, cat on features (:76-88)
                lead = [max(i, j) for i, j in zip(tn.shape[:-1], pn.shape[:-1])]
                joint = torch.cat((tn.expand(*lead, tn.shape[-1]), pn.expand(*lead, pn.shape[-1])), dim=-1)
            elif tn.dim() == 1:  # evaluation (:89-91)
                joint = torch.cat((tn, pn), dim=0)
            if self.joint_network is not None:
                joint = self.joint_network(joint)  # (:92-93)
        return self.nonlinearity(joint)  # (:95)
