"""CPU restatement of the projections either side of the joint (TEST ORACLE ONLY).

``speechbrain.nnet.linear.Linear.forward`` (vendor/speechbrain/speechbrain/nnet/linear.py:63-76) is ``nn.Linear`` on the
last dimension (:61,74), instantiated for ``encoder_proj`` / ``decoder_proj`` at
hparams/LibriSpeechMix/conformer-t_scratch.yaml:172-174,187-189 and called at train_librispeechmix_scratch.py:122,127;
its backward is autograd's: dX = dY W, dW = dY^T X, db = sum over rows of dY.  Restated in float64 numpy so that the fp32
reference result and the tcgen05 result can both be placed against the exact value.  Pinned against the reference's own
class by tests/golden/linear_proj*.npz (oracle/make_golden_projection.py).  Nothing in tsasr_b200/ imports this.
"""
import numpy as np


def linear_fwd(x, w, b=None):
    """y = x w^T + b on the last dimension (linear.py:74), float64.  x [..., K], w [N, K], b [N] or None."""
    y = np.asarray(x, dtype=np.float64) @ np.asarray(w, dtype=np.float64).T
    if b is not None:
        y = y + np.asarray(b, dtype=np.float64)
    return y


def linear_bwd(dy, x, w):
    """(dX, dW, db) of y = x w^T + b for 2-D x [R, K], dy [R, N], w [N, K], float64."""
    dy, x, w = (np.asarray(a, dtype=np.float64) for a in (dy, x, w))
    return dy @ w, dy.T @ x, dy.sum(axis=0)


def linear_abs_bound(x, w, b=None):
    """sum_k |x_k||w_k| (+|b|): the scale rounding errors of any finite-precision evaluation are proportional to."""
    s = np.abs(np.asarray(x, dtype=np.float64)) @ np.abs(np.asarray(w, dtype=np.float64)).T
    if b is not None:
        s = s + np.abs(np.asarray(b, dtype=np.float64))
    return s
