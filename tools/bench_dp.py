"""Time the wavefront DP kernel alone (development tool): python tools/bench_dp.py B T U"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsasr_b200 import ops
B, T, U = (int(x) for x in sys.argv[1:4])
dev = torch.device("cuda:0")
n = ops.lattice_elems(B, T, U)
lat2 = (-torch.rand(n, 2, device=dev) * 5 - 0.1)
ll = torch.full((B,), T, dtype=torch.int32, device=dev); tl = torch.full((B,), U - 1, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for warm in (True, False):
    ts = []
    for _ in range(6):
        if not warm: flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.alpha_beta(lat2, ll, tl, B, T, U); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts)//2]
    print(f"B={B} T={T} U={U} {'L2-warm' if warm else 'L2-cold'}: {t*1e3:.1f} us  {t*1e6/(T+U-1):.0f} ns/step  {24.0*B*T*U/t/1e6:.1f} GB/s")
