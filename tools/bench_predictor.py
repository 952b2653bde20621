"""Prediction network: the drop-in Embedding + LSTM against the reference's ops (one-hot
F.embedding -> pack_padded_sequence with the `.cpu()` of SB/nnet/RNN.py:35 -> cuDNN LSTM -> pad_packed_sequence) at the
recipe's shape, forward and forward+backward, CUDA-event timed (development tool).

    python tools/bench_predictor.py [--iters 20] [B U V Hd]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsasr_b200 import _lib  # noqa: E402
from tsasr_b200.predictor import Embedding, LSTM  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dims", nargs="*", type=int, default=[16, 100, 1000, 512])
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    B, U, V, Hd = a.dims
    d = torch.device("cuda:0")
    torch.manual_seed(0)
    emb = Embedding(num_embeddings=V, consider_as_one_hot=True, blank_id=0).to(d)
    ours = LSTM(input_shape=[None, None, V - 1], hidden_size=Hd).to(d)
    ref = torch.nn.LSTM(V - 1, Hd, batch_first=True).to(d)
    ref.load_state_dict({k[4:]: v for k, v in ours.state_dict().items()})
    tokens = torch.randint(1, V, (B, U), device=d)
    tokens[:, 0] = 0
    rel = (torch.rand(B, device=d) * 0.6 + 0.4)
    rel[0] = 1.0
    gy = torch.randn(B, U, Hd, device=d)

    def ours_fwd():
        return ours(emb(tokens), lengths=rel)[0]

    def ref_fwd():
        x = torch.nn.functional.embedding(tokens, emb.Embedding.weight, padding_idx=0)
        packed = torch.nn.utils.rnn.pack_padded_sequence(x, (rel * U).cpu(), batch_first=True, enforce_sorted=False)
        return torch.nn.utils.rnn.pad_packed_sequence(ref(packed)[0], batch_first=True)[0]

    def fb(f):
        f().backward(gy)

    out = {"shape": {"B": B, "U": U, "V": V, "hidden": Hd}}
    with torch.no_grad():
        out["fwd_us_ours"] = timeit(ours_fwd, a.iters)
        out["fwd_us_reference_ops"] = timeit(ref_fwd, a.iters)
        out["max_abs_diff_fwd"] = (ours_fwd() - ref_fwd()).abs().max().item()
    out["fwd_bwd_us_ours"] = timeit(lambda: fb(ours_fwd), a.iters)
    out["fwd_bwd_us_reference_ops"] = timeit(lambda: fb(ref_fwd), a.iters)
    _lib.kernel_timing(True)
    fb(ours_fwd)
    torch.cuda.synchronize()
    out["kernels_us"] = {k: round(v[0] * 1e3, 2) for k, v in _lib.kernel_timings().items()}
    _lib.kernel_timing(False)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
