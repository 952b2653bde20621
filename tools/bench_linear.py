"""Projection GEMMs: tsasr_b200.Linear vs the reference's op (torch fp32 nn.Linear, TF32 off) at the
recipe's shapes, forward and forward+backward, CUDA-event timed with an L2 flush between iterations (development tool).

    python tools/bench_linear.py [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsasr_b200  # noqa: E402
from tsasr_b200 import _lib  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
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
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    d = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
    out = {}
    for name, (R, K, N) in {"encoder_proj B16 T400": (6400, 256, 640), "decoder_proj B16 U100": (1600, 512, 640),
                            "encoder_proj B128 T400": (51200, 256, 640)}.items():
        ours, ref = tsasr_b200.Linear(N, input_size=K).to(d), torch.nn.Linear(K, N).to(d)
        x = torch.randn(R, K, device=d, requires_grad=True)
        gy = torch.randn(R, N, device=d)

        def fb(m):
            y = m(x)
            y.backward(gy)

        row = {"flop_fwd": 2.0 * R * K * N}
        with torch.no_grad():
            row["fwd_us_ours"] = timeit(lambda: ours(x), a.iters, flush)
            row["fwd_us_torch_fp32"] = timeit(lambda: ref(x), a.iters, flush)
        row["fwd_bwd_us_ours"] = timeit(lambda: fb(ours), a.iters, flush)
        row["fwd_bwd_us_torch_fp32"] = timeit(lambda: fb(ref), a.iters, flush)
        _lib.kernel_timing(True)
        fb(ours)
        torch.cuda.synchronize()
        row["kernels_us"] = {k: round(v[0] * 1e3, 2) for k, v in _lib.kernel_timings().items()}
        _lib.kernel_timing(False)
        out[name] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
