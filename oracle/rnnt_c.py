"""ctypes loader for oracle/rnnt_oracle.c (TEST ORACLE ONLY; see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librnnt_oracle.so")
_lib = None


def build(force=False):
    """Compile rnnt_oracle.c with the committed Makefile (gcc, pthreads)."""
    src = os.path.join(_HERE, "rnnt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        _lib.rnnt_oracle_torchaudio.argtypes = [fp, ip, ip, ip] + [ctypes.c_int] * 5 + [ctypes.c_float, ctypes.c_int, fp, fp]
        _lib.rnnt_oracle_torchaudio.restype = ctypes.c_int
        _lib.rnnt_oracle_numba.argtypes = [fp, ip, ip, ip] + [ctypes.c_int] * 6 + [fp, fp]
        _lib.rnnt_oracle_numba.restype = ctypes.c_int
        _lib.rnnt_oracle_num_threads.restype = ctypes.c_int
        _lib.rnnt_oracle_set_threads.argtypes = [ctypes.c_int]
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def num_threads():
    return lib().rnnt_oracle_num_threads()


def set_threads(n):
    lib().rnnt_oracle_set_threads(int(n))


def rnnt_torchaudio(logits, targets, logit_lengths, target_lengths, blank=0, clamp=-1.0, fp32=True, want_grads=True):
    """(costs[B], dlogits[B,T,U,V] or None) -- torchaudio semantics, see rnnt_oracle.c."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    B, T, U, V = logits.shape
    targets = np.ascontiguousarray(targets, dtype=np.int32).reshape(B, max(U - 1, 0))
    ll = np.ascontiguousarray(logit_lengths, dtype=np.int32)
    tl = np.ascontiguousarray(target_lengths, dtype=np.int32)
    costs = np.zeros(B, dtype=np.float32)
    grads = np.empty_like(logits) if want_grads else None
    rc = lib().rnnt_oracle_torchaudio(_f(logits), _i(targets), _i(ll), _i(tl), B, T, U, V, int(blank),
                                      float(clamp), int(bool(fp32)), _f(costs), _f(grads))
    if rc != 0:
        raise RuntimeError({-1: "blank must be within [0, logits.shape[-1])", -2: "input length mismatch",
                            -3: "output length mismatch"}.get(rc, "oracle error %d" % rc))
    return costs, grads


def rnnt_numba(log_probs, labels, T, U, blank=0, fp32=True, want_grads=True):
    """(per_utt[B] = -logP/T_b, grads w.r.t. log_probs or None) -- Numba semantics."""
    lp = np.ascontiguousarray(log_probs, dtype=np.float32)
    B, maxT, maxU, V = lp.shape
    labels = np.ascontiguousarray(labels, dtype=np.int32).reshape(B, max(maxU - 1, 0))
    T = np.ascontiguousarray(T, dtype=np.int32)
    U = np.ascontiguousarray(U, dtype=np.int32)
    per_utt = np.zeros(B, dtype=np.float32)
    grads = np.empty_like(lp) if want_grads else None
    rc = lib().rnnt_oracle_numba(_f(lp), _i(labels), _i(T), _i(U), B, maxT, maxU, V, int(blank),
                                 int(bool(fp32)), _f(per_utt), _f(grads))
    if rc != 0:
        raise RuntimeError("oracle error %d" % rc)
    return per_utt, grads
