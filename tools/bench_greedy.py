"""Whole greedy search: the reference's loop (per-utterance .item() reads every frame,
SB/decoders/transducer.py:138-218) against the on-device bookkeeping of tsasr_b200.decode, recipe-sized modules
(B=16 utterances, T=400 frames, H=640, V=1000, LSTM 512).  python tools/bench_greedy.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsasr_b200  # noqa: E402
from tsasr_b200 import decode  # noqa: E402
from oracle.greedy_decode import ToyPredictor, eager_joint_step, greedy_decode  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, T, V, E, HID, H = 16, 400, 1000, 128, 512, 640
pred = ToyPredictor(V, E, HID, H).to(dev).eval()
head = torch.nn.Linear(H, V).to(dev).eval()
with torch.no_grad():
    head.weight.mul_(4.0)
    head.bias[0] += 5.0      # blank wins on a good share of the frames, as in a trained model
tjoint = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
sm = torch.nn.LogSoftmax(dim=-1)
tn = 1.5 * torch.randn(B, T, H, device=dev)
eager = eager_joint_step(tjoint, [head], sm)
fused = decode.fused_joint_forward_step(tjoint, [head], sm)


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, out


rows = [("reference loop, eager joint step (per-frame .item() reads)", lambda: greedy_decode(tn, pred.layers(), eager)),
        ("reference loop, fused joint step (round 1)", lambda: greedy_decode(tn, pred.layers(), fused)),
        ("on-device bookkeeping, fused joint step", lambda: decode.greedy_decode_on_device(tn, pred.layers(), fused)),
        ("on-device bookkeeping, fused joint step, CUDA graph per frame", lambda: decode.greedy_decode_cuda_graph(tn, pred.layers(), fused))]
ref = None
print(f"greedy search of B={B} utterances x T={T} frames (H={H}, V={V}, LSTM {HID}); wall time of the whole search, best of 3")
for name, fn in rows:
    ms, out = timed(fn)
    hyps = out[0]
    if ref is None:
        ref = hyps
    print(f"  {name:64s} {ms:9.2f} ms   {ms / T * 1e3:7.1f} us/frame   hypotheses identical to the reference loop: {hyps == ref}"
          f"   (emitted {sum(len(h) for h in hyps)} labels)")
