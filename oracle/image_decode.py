"""Test-only views of the backward operand images and elementwise error bounds (TEST ORACLE INFRASTRUCTURE).

The fused backward never materialises dlogits [B,T,U,V] or the joint tensor [B,T,U,H]; it keeps them, chunk by chunk,
as bf16 SWIZZLE_128B operand images in its workspace.  The helpers here read those images back
into dense tensors so that tests can (1) compare dlogits element by element with the reference's, and (2) redo the
two backward GEMMs in float64 on exactly the operands the kernels consumed -- which separates GEMM / reduction /
scheduling bugs (must agree to fp32 accumulation accuracy) from operand rounding (bounded statistically below).
Single-chunk runs only: with several chunks the images of earlier chunks have been overwritten.
"""
import torch


def choose_tile_log2(T, U):
    """Same rule as choose_tile() in tsasr_b200/csrc/capi.cu: tT in {8,16,32}, least padded cells, 16 on ties."""
    best, best_l = -1, 4
    for l in range(3, 6):
        tT, tU = 1 << l, 128 >> l
        padded = ((T + tT - 1) // tT) * tT * ((U + tU - 1) // tU) * tU
        if best < 0 or padded < best or (padded == best and l == 4):
            best, best_l = padded, l
    return best_l


def tile_geometry(B, T, U):
    l = choose_tile_log2(T, U)
    tT, tU = 1 << l, 128 >> l
    nTt, nTu = (T + tT - 1) // tT, (U + tU - 1) // tU
    return tT, tU, nTt, nTu, B * nTt * nTu


def _unswizzle_index():
    # 16-byte chunk c of row r is stored at chunk c ^ (r & 7)  (sw128_offset, common.cuh)
    r = torch.arange(128)[:, None]
    c = torch.arange(8)[None, :]
    src_chunk = c ^ (r & 7)
    return (src_chunk[:, :, None] * 8 + torch.arange(8)[None, None, :]).reshape(128, 64)


def _images_to_dense(raw, B, T, U, cols):
    """raw [n_tiles, NB, 128, 64] (swizzled) -> [B, T, U, cols]; tile row r = ui * tT + ti."""
    tT, tU, nTt, nTu, n_tiles = tile_geometry(B, T, U)
    NB = raw.shape[1]
    idx = _unswizzle_index().to(raw.device)
    img = torch.gather(raw, 3, idx[None, None].expand(n_tiles, NB, 128, 64))
    rows = img.permute(0, 2, 1, 3).reshape(n_tiles, 128, NB * 64)[:, :, :cols]          # [tile, row, col]
    x = rows.reshape(B, nTt, nTu, tU, tT, cols).permute(0, 1, 4, 2, 3, 5)                # [B, tt, ti, tu, ui, col]
    return x.reshape(B, nTt * tT, nTu * tU, cols)[:, :T, :U].contiguous()


def decode_images(ws, B, T, U, H, V):
    """(dlogits [B,T,U,V], joint [B,T,U,H]) as float32 CPU tensors from the workspace of a SINGLE-chunk backward.
    Tiles the kernels never wrote (outside an utterance's T_b x U_b rectangle, or pruned) hold stale bytes: mask them
    with ``live_cell_mask`` / ``active_cell_mask`` before use."""
    _, _, _, _, n_tiles = tile_geometry(B, T, U)
    NT4 = 4 * ((V + 255) // 256)
    KB = H // 64
    base = (-ws.data_ptr()) % 1024
    dy_bytes = n_tiles * NT4 * 16384
    j_bytes = n_tiles * KB * 16384
    dy_raw = ws[base: base + dy_bytes].view(torch.bfloat16).view(n_tiles, NT4, 128, 64).float().cpu()
    j_raw = ws[base + dy_bytes: base + dy_bytes + j_bytes].view(torch.bfloat16).view(n_tiles, KB, 128, 64).float().cpu()
    return _images_to_dense(dy_raw, B, T, U, V), _images_to_dense(j_raw, B, T, U, H)


def live_cell_mask(B, T, U, logit_lengths, target_lengths):
    """[B,T,U] bool: cells inside the utterance's T_b x U_b rectangle."""
    t = torch.arange(T)[None, :, None]
    u = torch.arange(U)[None, None, :]
    return (t < logit_lengths.long()[:, None, None]) & (u < (target_lengths.long() + 1)[:, None, None])


def active_cell_mask(ws, stats_offset, B, T, U):
    """[B,T,U] bool: cells of the tiles the pruned backward kept (flags behind the statistics block of the workspace)."""
    tT, tU, nTt, nTu, n_tiles = tile_geometry(B, T, U)
    base = (-ws.data_ptr()) % 1024 + stats_offset + 64
    flags = ws[base: base + n_tiles].cpu().bool().reshape(B, nTt, 1, nTu, 1).expand(B, nTt, tT, nTu, tU)
    return flags.reshape(B, nTt * tT, nTu * tU)[:, :T, :U].contiguous()


def act_grad(enc, dec, act, act_param):
    """act'(enc + dec) as torch autograd and the kernels (backward_gemm.cuh, act_grad_pre) evaluate it: from the
    unrounded pre-activation (the bf16 rounding of the GEMM operand is a straight-through step)."""
    x = enc[:, :, None, :].double() + dec[:, None, :, :].double()
    if act == "leaky_relu":
        return torch.where(x > 0, torch.ones_like(x), torch.full_like(x, act_param))
    if act == "relu":
        return (x > 0).double()
    if act == "tanh":
        j = torch.tanh(x)
        return 1.0 - j * j
    return torch.ones_like(x)


def backward_from_operands(dlogits, joint, W, enc, dec, act, act_param, mask, dtype=torch.float64, with_bounds=True, with_sq=True):
    """float64 restatement of autograd through Linear (SB/nnet/linear.py:74), the activation and the broadcast add
    (SB/nnet/transducer/transducer_joint.py:74,95) on GIVEN dlogits / joint operands.  mask [B,T,U] zeroes cells that
    carry no data.  Returns dict(d_enc, d_dec, dW, db) plus the matching sums of |terms| ("abs_*") and of squared
    terms ("sq_*") used for elementwise error bounds."""
    m = mask[..., None].to(dtype)
    zero = torch.zeros((), dtype=dtype)
    dY = torch.where(mask[..., None], dlogits.to(dtype), zero)   # not a product: unwritten tiles may hold NaN / Inf bit patterns
    J = torch.where(mask[..., None], joint.to(dtype), zero)
    Wd = W.to(dtype)
    B, T, U, V = dY.shape
    H = J.shape[-1]
    dY2, J2 = dY.reshape(-1, V), J.reshape(-1, H)
    g = act_grad(enc, dec, act, act_param).to(dtype) * m                      # [B,T,U,H]
    dJ = (dY2 @ Wd).reshape(B, T, U, H) * g
    out = {"dW": dY2.t() @ J2, "db": dY2.sum(0), "d_enc": dJ.sum(2), "d_dec": dJ.sum(1)}
    del dJ
    if with_bounds:
        dYa = dY2.abs()
        out["abs_dW"] = dYa.t() @ J2.abs()
        out["abs_db"] = dYa.sum(0)
        dJ_abs = (dYa @ Wd.abs()).reshape(B, T, U, H) * g.abs()
        out["abs_d_enc"], out["abs_d_dec"] = dJ_abs.sum(2), dJ_abs.sum(1)
        del dJ_abs, dYa
    if with_bounds and with_sq:
        dYs = dY2 * dY2
        out["sq_dW"] = dYs.t() @ (J2 * J2)
        out["sq_db"] = dYs.sum(0)
        dJ_sq = (dYs @ (Wd * Wd)).reshape(B, T, U, H) * (g * g)
        out["sq_d_enc"], out["sq_d_dec"] = dJ_sq.sum(2), dJ_sq.sum(1)
    return out
