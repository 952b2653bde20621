"""Host-side cost of one drop-in training step (development tool): cProfile over the public-API path."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsasr_b200  # noqa: E402

B, T, U, V, H = 16, 400, 100, 1000, 640
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
enc = (0.5 * torch.randn(B, T, H, generator=g)).to(dev)
dec = (0.5 * torch.randn(B, U, H, generator=g)).to(dev)
head = torch.nn.Linear(H, V).to(dev)
tg = torch.randint(1, V, (B, U - 1), generator=g).to(dev)
il, tl = torch.ones(B, device=dev), torch.ones(B, device=dev)
joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)


def step(sync=True):
    e_, d_ = enc.detach().requires_grad_(), dec.detach().requires_grad_()
    t0 = time.perf_counter()
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
ts = [step() for _ in range(20)]
names = ["joiner+head", "transducer_loss (fwd launch)", "backward (launch)", "item()"]
for i, n in enumerate(names):
    print(f"{n:32s} {1e6 * sum(t[i] for t in ts) / len(ts):8.1f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
