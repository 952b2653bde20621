"""Whole beam search: the reference's searcher (utterances one after the other, one hypothesis
per network evaluation, SB/decoders/transducer.py:220-373) against tsasr_b200.decode.beam_search_batched (all utterances
concurrently, one batched evaluation + one device->host copy per round), recipe-sized modules (B=16, T=50 frames,
H=640, V=1000, LSTM 512, beam 4 and the recipe's 15).  Uses the real TransducerBeamSearcher when the reference install
(baseline/_ref) is present, else the coroutine search run one utterance at a time.  python tools/bench_beam.py"""
import os
import signal
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import tsasr_b200  # noqa: E402
from tsasr_b200 import decode  # noqa: E402
from oracle.greedy_decode import ToyPredictor, eager_joint_step  # noqa: E402

dev = torch.device(sys.argv[1] if len(sys.argv) > 1 else "cuda:0")   # "cpu": termination check of the toy model
torch.manual_seed(0)
B, T, V, E, HID, H = 16, 50, 1000, 128, 512, 640
pred = ToyPredictor(V, E, HID, H).to(dev).eval()
head = torch.nn.Linear(H, V).to(dev).eval()
with torch.no_grad():
    head.weight.mul_(2.0)
    head.bias[0] += 8.0      # blank wins on most frames, as in a trained model.  The reference's expansion loop has no limit on
                             # symbols per frame: on a random model with a weaker blank it does not terminate (checked on CPU)
tjoint = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
sm = torch.nn.LogSoftmax(dim=-1)
tn = 1.5 * torch.randn(B, T, H, device=dev)
fused = decode.fused_joint_forward_step(tjoint, [head], sm) if dev.type == "cuda" else eager_joint_step(tjoint, [head], sm)
eager = eager_joint_step(tjoint, [head], sm)

RefSearcher = None
try:
    import insitu_step

    if insitu_step.find_reference() is not None:
        insitu_step.import_reference()
        from speechbrain.decoders.transducer import TransducerBeamSearcher as RefSearcher
except Exception as ex:  # noqa: BLE001
    print("reference searcher not importable:", type(ex).__name__, ex)


class RowTimeout(Exception):
    pass


def _alarm(*_):
    raise RowTimeout()


signal.signal(signal.SIGALRM, _alarm)


def sync():
    if dev.type == "cuda":
        torch.cuda.synchronize()


def timed(fn, n=2, limit_s=60):
    """best-of-n wall time; a row that takes longer than limit_s is abandoned (never burn GPU minutes on a runaway search)"""
    signal.alarm(limit_s)
    try:
        fn()
        sync()
        best = 1e9
        for _ in range(n):
            t0 = time.perf_counter()
            out = fn()
            sync()
            best = min(best, time.perf_counter() - t0)
        return best * 1e3, out
    finally:
        signal.alarm(0)


for beam in (4, 15):
    cfg = dict(beam_size=beam, nbest=1, state_beam=2.3, expand_beam=2.3)
    rows = []
    if RefSearcher is not None:
        ref = RefSearcher(decode_network_lst=pred.layers(), tjoint=tjoint, classifier_network=[head], blank_id=0, **cfg)
        rows.append(("reference TransducerBeamSearcher (sequential, eager joint step)", lambda ref=ref: ref(tn)))
    rows += [("coroutine search, one utterance at a time, fused joint step",
              lambda: [decode.beam_search_batched(tn[i:i + 1], pred.layers(), fused, 0, **cfg)[0][0] for i in range(B)]),
             ("beam_search_batched: all utterances concurrently, eager joint step", lambda: decode.beam_search_batched(tn, pred.layers(), eager, 0, **cfg)),
             ("beam_search_batched: all utterances concurrently, fused joint step", lambda: decode.beam_search_batched(tn, pred.layers(), fused, 0, **cfg))]
    print(f"beam search, beam {beam}: B={B} utterances x T={T} frames (H={H}, V={V}, LSTM {HID}); wall time of the whole search, best of 2")
    first = None
    for name, fn in rows:
        try:
            ms, out = timed(fn)
        except RowTimeout:
            print(f"  {name:72s} abandoned after 60 s")
            continue
        hyps = out[0] if isinstance(out, tuple) else out
        if first is None:
            first = hyps
        print(f"  {name:72s} {ms:9.1f} ms   best hypotheses identical to the first row: {hyps == first}   (labels {sum(len(h) for h in hyps)})")
