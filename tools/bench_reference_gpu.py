"""The reference's own GPU path on this box, beside ours (evidence tool, SURVEY.md section 8d "Reference GPU path").

    python tools/bench_reference_gpu.py
Times, at BASELINE configs[1] (B=16, T=400, U=100, V=1000, H=640), one forward+backward of
  * the reference chain as the recipe runs it on a GPU: eager Transducer_joint("sum", LeakyReLU) + nn.Linear head +
    torchaudio.functional.rnnt_loss (CUDA kernels shipped in the wheel), fp32 and under bf16/fp16 autocast
    (torchaudio rejects bf16 logits, so the autocast run uses fp16 -- SURVEY 8a-L6);
  * tsasr_b200's drop-in path on the same tensors.
CUDA events, 3 warm-ups, median of 10, inputs resident in HBM."""
import os
import statistics
import sys

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
act = torch.nn.LeakyReLU()


def reference_step(autocast_dtype=None):
    from torchaudio.functional import rnnt_loss

    e_, d_ = enc.detach().requires_grad_(), dec.detach().requires_grad_()
    with torch.autocast("cuda", dtype=autocast_dtype, enabled=autocast_dtype is not None):
        joint = act(e_[..., None, :] + d_[:, None, ...])      # transducer_joint.py:73-74,95
        logits = head(joint)                                   # linear.py:74
    in_l = (il * logits.shape[1]).round().int()                # losses.py:58-59
    tg_l = (tl * tg.shape[1]).round().int()
    loss = rnnt_loss(logits if logits.dtype != torch.bfloat16 else logits.float(), tg.int(), in_l, tg_l, blank=0, reduction="mean")
    loss.backward()
    head.zero_grad(set_to_none=True)
    return loss


joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)


def ours_step():
    e_, d_ = enc.detach().requires_grad_(), dec.detach().requires_grad_()
    logits = head(joiner(e_[..., None, :], d_[:, None, ...]))
    loss = tsasr_b200.transducer_loss(logits, tg, il, tl, blank_index=0, reduction="mean", use_torchaudio=True)
    loss.backward()
    head.zero_grad(set_to_none=True)
    return loss


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts)


cells = B * T * U
rows = []
for name, fn in (("reference GPU path, fp32 (eager joint + Linear + torchaudio rnnt_loss CUDA)", lambda: reference_step(None)),
                 ("reference GPU path, fp16 autocast", lambda: reference_step(torch.float16)),
                 ("tsasr_b200 drop-in (fused tcgen05 path)", ours_step)):
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    l = fn()
    ms = timeit(fn)
    peak = (torch.cuda.max_memory_allocated() - base) / 2 ** 30
    rows.append((name, ms, peak, float(l)))
    print(f"{name:82s} {ms:8.2f} ms/step  {cells / ms / 1e3:8.1f} Mcells/s  peak extra memory {peak:6.2f} GiB  loss {float(l):.4f}")
print(f"speed-up over the reference GPU path: {rows[0][1] / rows[2][1]:.1f}x (fp32), {rows[1][1] / rows[2][1]:.1f}x (fp16 autocast)")
