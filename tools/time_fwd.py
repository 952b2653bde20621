"""Time the forward joint kernel (and optionally the GRAD pass via joint_bwd kernel timings) alone, L2 flushed
(development tool): python tools/time_fwd.py [B T U H V].  TSASR_B200_LIB selects the build (A/B runs)."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsasr_b200 import ops, _lib
a = [int(x) for x in sys.argv[1:6]]
B, T, U, H, V = a if len(a) == 5 else (16, 400, 100, 640, 1000)
dev = torch.device("cuda:0"); g = torch.Generator().manual_seed(0)
enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().to(dev); dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().to(dev)
W = ((torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5).bfloat16().to(dev); b = ((torch.rand(V, generator=g) * 2 - 1) / H ** 0.5).to(dev)
tg = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).to(dev)
ll = torch.full((B,), T, dtype=torch.int32).to(dev); tl = torch.full((B,), U - 1, dtype=torch.int32).to(dev)
dcost = torch.full((B,), 1.0 / B, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(25):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); lat2, logz = ops.joint_fwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01); e.record(); torch.cuda.synchronize()
    if i >= 5: ts.append(s.elapsed_time(e))
print(f"{os.environ.get('TSASR_B200_LIB', 'default')}: joint_fwd median {statistics.median(ts)*1e3:.1f} us  min {min(ts)*1e3:.1f} us  ({2.0*B*T*U*H*V/statistics.median(ts)/1e9:.0f} TFLOP/s)")
alpha, beta, cost, _, _ = ops.alpha_beta(lat2, ll, tl, B, T, U)
_lib.kernel_timing(True)
for i in range(8):
    flush.zero_()
    ops.joint_bwd(enc, dec, W, b, tg, ll, tl, 0, 0, 0.01, lat2, logz, alpha, beta, cost, dcost)
torch.cuda.synchronize()
print("  bwd kernels (us):", {k: round(v[0] / 8 * 1e3, 1) for k, v in _lib.kernel_timings().items()})
