"""How accurate is the REFERENCE itself?  torchaudio's CPU rnnt_loss (fp32 log-domain DP, the op behind
SB/nnet/losses.py:72-79) against the float64 restatement (oracle/rnnt_numpy.py) on lattices whose total log-likelihood
has the magnitude of BASELINE configs[1] / configs[3] (|L| ~ 3.5e3 / 8e3: ~log V nats per step over T+U steps).
fp32 alpha/beta carry ~|L| * 2^-24 absolute error per rounding, exp(alpha + beta - L) turns it into a RELATIVE error of
the whole cell's gradient row -- common to every implementation that keeps the lattice in fp32 (the reference does,
so do we).  The tolerances of tests/test_parity_tight_gpu.py carry that floor explicitly.  CPU only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rnnt_numpy as rn  # noqa: E402
from oracle.reference_chain import reference_rnnt_abs  # noqa: E402

for T, U, logV in ((24, 9, 3.7), (400, 100, 6.9), (750, 200, 8.5)):
    rng = np.random.default_rng(0)
    V = 8
    logits = rng.standard_normal((1, T, U, V)).astype(np.float32)
    logits[..., V - 1] += np.float32(logV + 2.0)  # a junk class soaks up the mass: the other log-probs sit near -logV
    targets = rng.integers(1, V - 1, (1, U - 1)).astype(np.int32)
    ll, tl = np.array([T], np.int32), np.array([U - 1], np.int32)
    c64, g64 = rn.rnnt_torchaudio(logits, targets, ll, tl, 0)
    x = torch.tensor(logits, requires_grad=True)
    c32 = reference_rnnt_abs(x, torch.tensor(targets), torch.tensor(ll), torch.tensor(tl), 0, "none")
    c32.sum().backward()
    g32 = x.grad.numpy().astype(np.float64)
    row = np.abs(g64).sum(-1)  # per-cell scale of the gradient row
    keep = row > 1e-6 * row.max()
    rel = (np.abs(g32 - g64).sum(-1) / np.maximum(row, 1e-300))[keep]
    model = 2.0 ** -23 * abs(c64[0]) * np.sqrt(T + U)
    print(f"T={T} U={U} |L|={abs(c64[0]):8.1f}: loss rel err {abs(c32.item() - c64[0]) / abs(c64[0]):.1e};  per-cell relative error of the "
          f"fp32 reference's dlogits row: median {np.median(rel):.1e}  99% {np.quantile(rel, 0.99):.1e}  max {rel.max():.1e}"
          f"   [2^-23 |L| sqrt(T+U) = {model:.1e}]")
