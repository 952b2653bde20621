"""ctypes binding of libtsasr_b200.so (the C ABI declared in include/tsasr_b200.h).

There is NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

from . import _build

_lib = None

ABI_VERSION = 4  # TSASR_ABI_VERSION of include/tsasr_b200.h
E_INVALID, E_UNSUPPORTED, E_CUDA, E_WORKSPACE = -1, -2, -3, -4
F32, F16, BF16 = 0, 1, 2
ACT_CODES = {"leaky_relu": 0, "relu": 1, "tanh": 2, "identity": 3}

_vp, _i, _f, _sz, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/tsasr_b200.h one to one
SIGNATURES = {
    "tsasr_abi_version": (_i, []),
    "tsasr_last_error": (ctypes.c_char_p, []),
    "tsasr_launch_count": (_ll, []),
    "tsasr_lattice_elems": (_sz, [_i, _i, _i]),
    "tsasr_logits_to_lattice": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "tsasr_lattice_alpha_beta": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsasr_logits_grad": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp]),
    "tsasr_logprobs_grad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsasr_joint_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "tsasr_joint_loss_fwd_layout": (_i, [_i, _i, _i, _i, _i, _vp]),
    "tsasr_joint_loss_fwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _sz,
                                  _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsasr_linear_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "tsasr_linear_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "tsasr_linear_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "tsasr_joint_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _ll]),
    "tsasr_joint_bwd_stats_offset": (_sz, [_i, _i, _i, _i, _i, _ll]),
    "tsasr_joint_bwd": (_i, [_vp] * 7 + [_i] * 7 + [_f] + [_vp] * 7 + [_sz, _ll, _f, _f] + [_vp] * 5),
    "tsasr_prepare_lengths": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "tsasr_cast_operands_bf16": (_i, [_vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp]),
    "tsasr_lstm_fwd": (_i, [_vp, _i, _i, _i] + [_vp] * 7 + [_i, _i, _i] + [_vp] * 7 + [_vp]),
    "tsasr_lstm_bwd": (_i, [_vp] * 7 + [_i, _i, _i, _vp, _vp]),
    "tsasr_onehot_dw": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "tsasr_lstm_fwd": (_i, [_vp, _i, _i, _i] + [_vp] * 7 + [_i, _i, _i] + [_vp] * 7 + [_vp]),
    "tsasr_lstm_bwd": (_i, [_vp] * 7 + [_i, _i, _i, _vp, _vp]),
    "tsasr_onehot_dw": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "tsasr_joint_decode_workspace_bytes": (_sz, [_i]),
    "tsasr_joint_decode_step": (_i, [_vp, _vp, _ll, _ll, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _sz, _vp]),
    "tsasr_kernel_timing_enable": (_i, [_i]),
    "tsasr_kernel_timings": (_i, [ctypes.c_char_p, _vp, _vp, _i]),
    "tsasr_joint_debug_logits": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp]),
}


def so_path():
    return _build.SO_PATH


def load():
    """Load the shared library (never builds implicitly on the GPU box: the .so travels in-tree)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.SO_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: the CUDA extension has not been built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). tsasr_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.tsasr_abi_version() != ABI_VERSION:
        raise RuntimeError("libtsasr_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().tsasr_last_error().decode("utf-8", "replace")
        exc = {E_INVALID: ValueError, E_UNSUPPORTED: NotImplementedError}.get(rc, RuntimeError)
        raise exc(f"tsasr_b200 error {rc}: {msg}")


def launch_count():
    return int(load().tsasr_launch_count())


def kernel_timing(on):
    """Bracket every kernel launch of the library with CUDA events (measurement aid, no host sync)."""
    load().tsasr_kernel_timing_enable(1 if on else 0)


def kernel_timings(max_n=32):
    """-> {kernel name: (total ms, launches)} since the last call; waits for the recorded events."""
    names = ctypes.create_string_buffer(32 * max_n)
    ms = (ctypes.c_float * max_n)()
    counts = (ctypes.c_int * max_n)()
    n = load().tsasr_kernel_timings(names, ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(counts, ctypes.c_void_p), max_n)
    out = {}
    for i in range(n):
        out[names.raw[32 * i: 32 * i + 32].split(b"\0")[0].decode()] = (float(ms[i]), int(counts[i]))
    return out
