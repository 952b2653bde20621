"""Development tool (GPU box): error statistics of the fused backward against (a) the CPU reference chain and (b) a
float64 redo of the backward GEMMs on the decoded operand images.  Prints the numbers the tolerances in
tests/test_parity_tight_gpu.py are chosen from."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsasr_b200  # noqa: E402
from tsasr_b200 import _lib, ops  # noqa: E402
from oracle import image_decode as imd  # noqa: E402
from oracle.reference_chain import reference_joint_loss_fwd_bwd  # noqa: E402


def inputs(B, T, U, H, V, seed, ragged=True):
    gen = torch.Generator().manual_seed(seed)
    enc = (0.5 * torch.randn(B, T, H, generator=gen)).bfloat16()
    dec = (0.5 * torch.randn(B, U, H, generator=gen)).bfloat16()
    bound = 1.0 / (H ** 0.5)
    W = ((torch.rand(V, H, generator=gen) * 2 - 1) * bound).bfloat16()
    b = (torch.rand(V, generator=gen) * 2 - 1) * bound
    targets = torch.randint(1, V, (B, max(U - 1, 0)), generator=gen, dtype=torch.int32)
    ll = torch.full((B,), T, dtype=torch.int32)
    tl = torch.full((B,), U - 1, dtype=torch.int32)
    if ragged and B > 1:
        ll[1:] = torch.randint(max(1, T // 2), T + 1, (B - 1,), generator=gen, dtype=torch.int32)
        tl[1:] = torch.randint(0, U, (B - 1,), generator=gen, dtype=torch.int32)
    return enc, dec, W, b, targets, ll, tl


def main():
    d = torch.device("cuda:0")
    shapes = [((2, 24, 9, 64, 40), "tanh"), ((2, 40, 17, 640, 1000), "leaky_relu"), ((3, 20, 1, 128, 50), "leaky_relu"),
              ((2, 30, 40, 320, 29), "relu"), ((3, 300, 80, 128, 500), "leaky_relu"), ((2, 400, 100, 640, 1000), "leaky_relu")]
    for (B, T, U, H, V), act in shapes:
        enc, dec, W, b, targets, ll, tl = inputs(B, T, U, H, V, seed=B + T + U + H + V)
        dcost = torch.linspace(0.5, 1.5, B)
        ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, act, 0.01, round_bf16=True, dcost=dcost,
                                           keep_intermediates=True)
        live = imd.live_cell_mask(B, T, U, ll, tl)
        stat = imd.backward_from_operands(ref["dlogits"], ref["joint"], W.float(), enc.float(), dec.float(), act, 0.01, live)
        for eps in (0.0, -30.0):
            e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
            costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, activation=act,
                                                     reduction="none", max_chunk_cells=1 << 40, prune_log2_eps=eps)
            (costs * dcost.to(d)).sum().backward()
            torch.cuda.synchronize()
            got = {"d_enc": e.grad.cpu(), "d_dec": dc.grad.cpu(), "dW": w.grad.cpu(), "db": bb.grad.cpu()}
            ws = ops.last_workspace(d)
            dY, J = imd.decode_images(ws, B, T, U, H, V)
            mask = live
            if eps < 0:
                off = int(_lib.load().tsasr_joint_bwd_stats_offset(B, T, U, H, V, 1 << 40))
                mask = imd.active_cell_mask(ws, off, B, T, U) & live
            exact = imd.backward_from_operands(dY, J, W.float(), enc.float(), dec.float(), act, 0.01, mask)
            print(f"--- B={B} T={T} U={U} H={H} V={V} act={act} prune={eps} active cells {int(mask.sum())}/{int(live.sum())}"
                  f"  loss rel err {((costs.detach().cpu() - ref['costs']).abs() / ref['costs'].abs()).max().item():.2e}")
            dl_err = ((dY - ref["dlogits"]).abs() * mask[..., None]).max().item()
            dl_rel = (((dY - ref["dlogits"]).abs() - ref["dlogits"].abs() * 2.0 ** -8) * mask[..., None]).max().item()
            pruned_max = (ref["dlogits"].abs() * (live & ~mask)[..., None]).max().item()
            jerr = ((J - ref["joint"]).abs() * mask[..., None]).max().item()
            print(f"    dlogits images: max abs err {dl_err:.2e}, max (err - 2^-8 |ref|) {dl_rel:.2e}; largest reference |dlogits| in pruned cells"
                  f" {pruned_max:.2e}; J image max err {jerr:.2e}")
            for k in ("d_enc", "d_dec", "dW", "db"):
                g, r, x = got[k].double(), ref[k].double(), exact[k]
                err_ref = (g - r).abs()
                err_x = (g - x).abs()
                sigma = stat["sq_" + k].sqrt() * 2.0 ** -9 / 3 ** 0.5
                rms = r.pow(2).mean().sqrt().item()
                print(f"    {k:6s} vs ref: max/max {err_ref.max().item() / r.abs().max().item():.2e}  max err/(3e-3|ref|+3e-3 rms) "
                      f"{(err_ref / (3e-3 * r.abs() + 3e-3 * rms)).max().item():.2f}  max err/sigma {(err_ref / (sigma + 1e-30)).max().item():.1f}"
                      f"  max err/abs-sum {(err_ref / (stat['abs_' + k] + 1e-30)).max().item():.2e}"
                      f" | vs fp64(images): max err/abs-sum {(err_x / (exact['abs_' + k] + 1e-30)).max().item():.2e}  max/max "
                      f"{err_x.max().item() / x.abs().max().item():.2e}")


def config4():
    """Config-4 width (T=750, U=200, V=5000, H=640): default chunking at B=2 (two chunks, accumulate=1 in the second,
    k=3 dW schedule) and three forced chunks at B=1, against the CPU reference chain."""
    import time
    d = torch.device("cuda:0")
    for B, chunk, ragged in ((1, 128 * 400, False), (2, 0, True)):
        T, U, H, V = 750, 200, 640, 5000
        enc, dec, W, b, targets, ll, tl = inputs(B, T, U, H, V, seed=40 + B, ragged=ragged)
        if ragged:
            ll[1], tl[1] = 533, 140
        dcost = torch.linspace(0.5, 1.5, B)
        e, dc, w, bb = (x.to(d).float().requires_grad_() for x in (enc, dec, W, b))
        costs = tsasr_b200.fused_joint_rnnt_loss(e, dc, w, bb, targets.to(d), ll.to(d), tl.to(d), blank=0, reduction="none",
                                                 max_chunk_cells=chunk)
        (costs * dcost.to(d)).sum().backward()
        torch.cuda.synchronize()
        got = {"d_enc": e.grad.cpu(), "d_dec": dc.grad.cpu(), "dW": w.grad.cpu(), "db": bb.grad.cpu()}
        t0 = time.time()
        ref = reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=True, dcost=dcost,
                                           keep_intermediates=True)
        t1 = time.time()
        live = imd.live_cell_mask(B, T, U, ll, tl)
        stat = imd.backward_from_operands(ref["dlogits"], ref["joint"], W.float(), enc.float(), dec.float(), "leaky_relu", 0.01, live,
                                          dtype=torch.float32)
        t2 = time.time()
        print(f"--- config-4 width B={B} chunk={chunk}: reference chain {t1 - t0:.1f} s, bounds {t2 - t1:.1f} s; active/live tiles "
              f"{ops.last_backward_tile_stats(d)}; loss rel err {((costs.detach().cpu() - ref['costs']).abs() / ref['costs'].abs()).max().item():.2e}")
        for k in ("d_enc", "d_dec", "dW", "db"):
            g, r = got[k].double(), ref[k].double()
            err = (g - r).abs()
            sigma = stat["sq_" + k].double().sqrt() * 2.0 ** -9 / 3 ** 0.5
            rms = r.pow(2).mean().sqrt().item()
            print(f"    {k:6s} vs ref: max/max {err.max().item() / r.abs().max().item():.2e}  max err/(3e-3|ref|+3e-3 rms) "
                  f"{(err / (3e-3 * r.abs() + 3e-3 * rms)).max().item():.2f}  max err/sigma {(err / (sigma + 1e-30)).max().item():.1f}"
                  f"  max err/abs-sum {(err / (stat['abs_' + k].double() + 1e-30)).max().item():.2e}  fp32-oracle self check "
                  f"{((stat[k].double() - r).abs() / (stat['abs_' + k].double() + 1e-30)).max().item():.2e}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "config4":
        config4()
    else:
        main()
