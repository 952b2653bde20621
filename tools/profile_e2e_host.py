"""Host-side timeline of one drop-in training step (development tool): when, relative to the start of the step, does each
C-ABI call of the library get issued, and how long does the host spend in each section of the public-API path?
The time before the first GEMM launch is GPU idle time in a loop that reads its loss every step."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsasr_b200  # noqa: E402
from tsasr_b200 import _lib  # noqa: E402

B, T, U, V, H = 16, 400, 100, 1000, 640
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
enc = (0.5 * torch.randn(B, T, H, generator=g)).to(dev)
dec = (0.5 * torch.randn(B, U, H, generator=g)).to(dev)
head = torch.nn.Linear(H, V).to(dev)
tg = torch.randint(1, V, (B, U - 1), generator=g).to(dev)
il, tl = torch.ones(B, device=dev), torch.ones(B, device=dev)
joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)

# ---- wrap every C entry point with a timestamp recorder ----
lib = _lib.load()
t_step0 = [0.0]
calls = []


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        calls.append((self.name, 1e6 * (t0 - t_step0[0]), 1e6 * (time.perf_counter() - t0)))
        return r


class TimedLib:
    def __init__(self, lib):
        self._lib = lib
        self._cache = {}

    def __getattr__(self, name):
        if name not in self._cache:
            self._cache[name] = Timed(name, getattr(self._lib, name))
        return self._cache[name]


_lib._lib = TimedLib(lib)


def step(sync=True):
    e_, d_ = enc.detach().requires_grad_(), dec.detach().requires_grad_()
    t_step0[0] = t0 = time.perf_counter()
    logits = head(joiner(e_[..., None, :], d_[:, None, ...]))
    t1 = time.perf_counter()
    loss = tsasr_b200.transducer_loss(logits, tg, il, tl, blank_index=0, reduction="mean", use_torchaudio=True)
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    head.zero_grad(set_to_none=True)
    v = loss.item() if sync else None
    t4 = time.perf_counter()
    return (t1 - t0, t2 - t1, t3 - t2, t4 - t3)


for _ in range(5):
    step()
calls.clear()
ts = [step() for _ in range(20)]
names = ["joiner+head", "transducer_loss (fwd launch)", "backward (launch)", "item()"]
for i, n in enumerate(names):
    print(f"{n:32s} {1e6 * sum(t[i] for t in ts) / len(ts):8.1f} us")
print("C-ABI calls of one step: issue time after the start of the step / host time inside the call (us, mean of 20 steps)")
order, agg = [], {}
for name, at, dur in calls:
    if name not in agg:
        order.append(name)
        agg[name] = [0.0, 0.0, 0]
    agg[name][0] += at
    agg[name][1] += dur
    agg[name][2] += 1
for name in order:
    a = agg[name]
    print(f"  {name:36s} at {a[0] / a[2]:8.1f}   inside {a[1] / a[2]:7.1f}   ({a[2] // 20} per step)")
_lib._lib = lib
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
s.record()
for _ in range(20):
    step()
e.record()
torch.cuda.synchronize()
print(f"steps back to back with a loss read each: {s.elapsed_time(e) / 20:.3f} ms per step")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
