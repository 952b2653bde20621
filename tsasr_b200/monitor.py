"""Loss-side conveniences around the hot path: the non-finite check without a host sync.

``Brain.fit_batch`` calls ``self.check_gradients(loss)`` between the forward and the backward pass
(vendor/speechbrain/speechbrain/core.py:1072); its first statement, ``if loss.isfinite():`` (:1130), blocks the host until
the whole forward pass and the loss have finished on the GPU -- and only then does the host start queueing the backward
pass, so the GPU idles while a few thousand backward kernels are being launched.  ``fit_batch`` ignores the return value
of the check; what the check DOES is bookkeeping for the rare non-finite case (count, warnings, ``ValueError`` once
``nonfinite_patience`` is exhausted, :1133-1150).

``install(brain)`` replaces ``brain.check_gradients`` by a deferred version: the finiteness flag of this step's loss is
computed on the device and copied to pinned host memory asynchronously; it is LOOKED AT when the next step's check runs
(by then it has long arrived), and only a non-finite flag runs the reference's own ``check_gradients`` on that loss --
same counters, same warnings, same exception, one step later.  ``uninstall`` restores the original; ``flush`` resolves
the pending flag (end of an epoch).  Nothing here touches the kernels; CPU tensors fall through to the original check.
"""
import torch


class DeferredFiniteCheck:
    def __init__(self, brain):
        self.brain = brain
        self.original = brain.check_gradients
        self.pending = None
        self.deferred_steps = 0

    def _resolve(self):
        if self.pending is None:
            return True
        host, event, loss = self.pending
        self.pending = None
        event.synchronize()  # recorded a whole step ago: returns immediately
        if bool(host[0]):
            return True
        return self.original(loss)  # the reference's bookkeeping for a non-finite loss (counts, warns, raises on patience)

    def __call__(self, loss):
        if not (isinstance(loss, torch.Tensor) and loss.is_cuda):
            return self.original(loss)
        ok = self._resolve()  # last step's flag
        flag = torch.isfinite(loss.detach()).reshape(1)
        host = torch.empty((1,), dtype=torch.bool).pin_memory()
        host.copy_(flag, non_blocking=True)
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(loss.device))
        self.pending = (host, event, loss.detach())
        self.deferred_steps += 1
        return ok

    def flush(self):
        """Resolve the pending flag now (synchronises with the step that produced it)."""
        return self._resolve()


def install(brain):
    """Replace ``brain.check_gradients`` (SB/core.py:1115-1150) by the deferred, sync-free check; returns the checker."""
    chk = DeferredFiniteCheck(brain)
    brain.check_gradients = chk
    return chk


def uninstall(brain):
    chk = brain.check_gradients
    if isinstance(chk, DeferredFiniteCheck):
        chk.flush()
        brain.check_gradients = chk.original
