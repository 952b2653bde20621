"""One forward + one backward of the fused joint + loss at a given shape (profiling target):
python tools/one_step.py [B T U H V] [reps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsasr_b200 import ops
a = [int(x) for x in sys.argv[1:6]]
B, T, U, H, V = a if len(a) == 5 else (16, 400, 100, 640, 1000)
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
dev = torch.device("cuda:0"); g = torch.Generator().manual_seed(0)
enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().to(dev); dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().to(dev)
W = ((torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5).bfloat16().to(dev); b = ((torch.rand(V, generator=g) * 2 - 1) / H ** 0.5).to(dev)
tg = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).to(dev)
ll = torch.full((B,), T, dtype=torch.int32).to(dev); tl = torch.full((B,), U - 1, dtype=torch.int32).to(dev)
dcost = torch.full((B,), 1.0 / B, device=dev)
for _ in range(reps):
    lat2, logz = ops.joint_fwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01)
    alpha, beta, cost, _, _ = ops.alpha_beta(lat2, ll, tl, B, T, U)
    ops.joint_bwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01, lat2, logz, alpha, beta, cost, dcost)
torch.cuda.synchronize()
print("ok", float(cost.sum()))
