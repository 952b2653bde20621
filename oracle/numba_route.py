"""Route 2 of SURVEY.md section 8c: the reference's own Numba kernels, run on CPU (TEST ORACLE ONLY).

Loads /root/reference/vendor/speechbrain/speechbrain/nnet/loss/transducer_loss.py BY FILE PATH
under NUMBA_ENABLE_CUDASIM=1.  Only usable in the authoring container (the GPU box has no
/root/reference); used by make_golden.py to produce tests/golden/*.npz.  Very slow (Python-thread
simulator): tiny shapes only.
"""
import importlib.util
import os

REF_FILE = "/root/reference/vendor/speechbrain/speechbrain/nnet/loss/transducer_loss.py"


def available():
    return os.path.exists(REF_FILE)


def load():
    os.environ["NUMBA_ENABLE_CUDASIM"] = "1"  # must be set before numba is imported
    spec = importlib.util.spec_from_file_location("ref_transducer_loss", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def numba_loss_and_grad(logits, labels, T, U, blank, reduction):
    """Transducer.apply(logits.log_softmax(-1), ...) as SB/nnet/losses.py:84-87 calls it.

    Returns (loss, d loss / d logits through autograd, stored grads w.r.t. log-probs)."""
    import torch

    mod = load()
    logits = logits.detach().clone().float().requires_grad_()
    log_probs = logits.log_softmax(-1)
    log_probs.retain_grad()
    loss = mod.Transducer.apply(log_probs, labels.int(), T.int(), U.int(), blank, reduction)
    (loss.sum() if loss.dim() else loss).backward()
    return loss.detach(), logits.grad.detach(), log_probs.grad.detach()
