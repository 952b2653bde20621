"""One projection forward+backward at a chosen shape (target command of ncu captures of linear_gemm_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsasr_b200  # noqa: E402

R, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (51200, 256, 640)
d = torch.device("cuda:0")
m = tsasr_b200.Linear(N, input_size=K).to(d)
x = torch.randn(R, K, device=d, requires_grad=True)
gy = torch.randn(R, N, device=d)
for _ in range(3):
    y = m(x)
    y.backward(gy)
torch.cuda.synchronize()
