// backward_gemm.cuh -- the two backward GEMMs of the fused joint (tcgen05 + TMEM + TMA/bulk copies).
//
// They consume the operand images the MODE_GRAD pass of joint_gemm.cuh leaves in the (L2-sized)
// chunk workspace: per 128-cell tile, dY images [v-block][128 cells x 64 v] and J images
// [h-block][128 cells x 64 h], both bf16 in the SWIZZLE_128B shared-memory image, so a block is one
// contiguous 16 KB bulk copy and is directly addressable by a UMMA descriptor -- K-major when the
// 64-wide dimension is the contraction (dJ), MN-major when the 128 cells are the contraction (dW).
//
//   dj_gemm_kernel : dJ[cells, h] = sum_v dY[cells, v] * W[v, h]          (autograd of SB/nnet/linear.py:74)
//                    epilogue: dpre = dJ * act'(J) (transducer_joint.py:95 backward) and the two
//                    broadcast-sum reductions of transducer_joint.py:74 (sum over u -> d_enc rows,
//                    sum over t -> d_dec rows) inside the tile; per-tile partial rows go to a small
//                    buffer that reduce_dpre_* kernels fold deterministically.
//   dw_gemm_kernel : dW[v, h] += sum_cells dY[cells, v] * J[cells, h], db[v] += sum_cells dY[cells, v]
//                    (db rides along as 16 extra accumulator columns fed by a constant "ones" block).
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "joint_gemm.cuh"

namespace tsasr {

static constexpr int kImgBytes = 16384;  // one [128 x 64] bf16 SWIZZLE_128B image
static constexpr int kBwdThreads = 192;  // dW kernel: warps 0-3: epilogue, warp 4: loads, warp 5: MMA (highest id = highest issue priority)
static constexpr int kBwdWarpLoad = 4;
static constexpr int kBwdWarpMma = 5;

struct BwdParams {
    const int* logit_lengths;
    const int* target_lengths;
    int B, T, U, H, V;
    int act_kind;
    float act_param;
    int tT_log2, nTt, nTu;
    int tile_begin, tile_end;
    int KB;        // H / 64
    int NVB;       // ceil(V / 64): v-blocks that carry data
    int NT4;       // v-block images allocated per tile (4 * ceil(V / 256))
    const __nv_bfloat16* dY_img;
    const __nv_bfloat16* J_img;
    // dJ
    const __nv_bfloat16* enc;  // [B,T,H]
    const __nv_bfloat16* dec;  // [B,U,H]
    float* dpre_part;  // [tile - tile_begin][tT + tU][H]
    int NHC;           // ceil(H / 256): chunks of h-rows per tile pair
    // dW
    float* dW_part;    // [n_splits][NVT * 128][H]
    float* db_part;    // [n_splits][NVT * 128]
    int NVT;           // ceil(V / 128)
    int NHT;           // h-tiles of up to 4 h-blocks
    int n_splits;
    int accumulate;    // 0: store, 1: read-modify-write (later chunks)
    long long* prof;   // development: MMA-warp cycle counters per CTA (or nullptr)
};

__device__ __forceinline__ bool tile_live(const BwdParams& p, int tile) {
    const int per_b = p.nTt * p.nTu;
    const int b = tile / per_b;
    const int rem = tile - b * per_b;
    const int tt = rem / p.nTu, tu = rem - tt * p.nTu;
    return (tt << p.tT_log2) < p.logit_lengths[b] && tu * (kTileM >> p.tT_log2) < p.target_lengths[b] + 1;
}

// ------------------------------------------------------------------------------------------------
// dJ GEMM (transposed) + activation backward + in-register broadcast-sum reductions
// ------------------------------------------------------------------------------------------------
// The GEMM is computed TRANSPOSED, dJ^T[h, cell] = sum_v W[v, h] * dY[cell, v], on CTA pairs
// (tcgen05 cta_group::2, M = 256 h-rows x N = 256 cells = two cell tiles, K = v):
//   A = W^T   MN-major: per k-block two TMA boxes [64 v x 64 h] SWIZZLE_128B per CTA (its own 128 h-rows)
//   B = dY    K-major : the [128 cells x 64 v] operand image of the CTA's own cell tile (tile 2i + rank)
// so a TMEM lane is one h and the 128 columns of a tile are its cells.  Both broadcast-sum reductions
// of transducer_joint.py:74 (sum over u -> d_enc row, sum over t -> d_dec row) are then plain
// in-register adds along the columns of a thread, act'(enc[t,h] + dec[u,h]) needs 24 scalars per thread,
// and the per-tile partial rows are stored h-contiguous (one 128-byte line per warp store).
// Unit of work = (tile pair, chunk of 256 h-rows); H = 640 runs three chunks, the upper half of the last
// one is padding (its CTA only supplies its dY image).  Accumulators are double-buffered (2 x 256 TMEM
// columns), so the epilogue of a unit overlaps the MMAs of the next.
static constexpr int kDjStages = 5;
static constexpr int kDjStageBytes = 2 * 8192 + kImgBytes;  // two W boxes + one dY image
static constexpr int kDjThreads = 320;  // warps 0-7: epilogue, warp 8: loads, warp 9: MMA (leader CTA)
static constexpr int kDjWarpLoad = 8;
static constexpr int kDjWarpMma = 9;
static constexpr int kDjChunkH = 256;

struct DjSmem { uint32_t stage_off, bar_off, tmem_off, total; };
__host__ __device__ inline DjSmem dj_smem_layout() {
    DjSmem l;
    l.stage_off = 0;
    l.bar_off = kDjStages * kDjStageBytes;
    l.tmem_off = l.bar_off + 16 * 8;
    l.total = l.tmem_off + 16;
    return l;
}

// four K = 16 steps of one stage: A (MN-major) advances by 16 v-rows = 2048 B, B (K-major) by 32 B
__device__ __forceinline__ void umma_bf16_2cta_x4_mnA_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                        uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "add.u64 a1, %1, 128;\n\tadd.u64 a2, %1, 256;\n\tadd.u64 a3, %1, 384;\n\t"
        "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, t;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, t;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, t;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate_first) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta_e(void* smem_dst, const void* tmap, uint32_t mbar_cluster, int x, int y) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(mbar_cluster), "r"(x), "r"(y) : "memory");
}

// act'(x) of the pre-activation x = enc + dec, evaluated exactly as the forward does it: the derivative is
// taken at the bf16-rounded activation output (sign for (leaky) ReLU, 1 - J^2 for tanh).
template <int ACT>
__device__ __forceinline__ float act_grad_pre(float x, float param) {
    if (ACT == ACT_LEAKY_RELU) return x > 0.f ? 1.f : param;
    if (ACT == ACT_RELU) return x > 0.f ? 1.f : 0.f;
    if (ACT == ACT_TANH) {
        const float j = __bfloat162float(__float2bfloat16_rn(tanhf(x)));
        return 1.f - j * j;
    }
    return 1.f;
}

// Epilogue of one (tile, 32 h-rows) slab: thread = one h, registers = the tile's 128 cells.
template <int TT_LOG2, int ACT>
__device__ __forceinline__ void dj_epilogue_tile(const BwdParams& p, uint32_t tmem_addr, const float (&e)[1 << TT_LOG2],
                                                 const float (&d)[128 >> TT_LOG2], float* __restrict__ out_h) {
    constexpr int TT = 1 << TT_LOG2, TU = 128 >> TT_LOG2;
    float acc_e[TT], acc_d[TU];
#pragma unroll
    for (int i = 0; i < TT; ++i) acc_e[i] = 0.f;
#pragma unroll
    for (int i = 0; i < TU; ++i) acc_d[i] = 0.f;
    uint32_t raw0[32], raw1[32];
    auto consume = [&](const uint32_t (&raw)[32], const int j) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int r = 32 * j + i;  // tile row = ui * TT + ti
            const int ti = r & (TT - 1), ui = r >> TT_LOG2;
            const float v = __uint_as_float(raw[i]) * act_grad_pre<ACT>(e[ti] + d[ui], p.act_param);
            acc_e[ti] += v;
            acc_d[ui] += v;
        }
    };
    tmem_ld_32x32b_x32(tmem_addr, raw0);
    tmem_ld_wait();
    tmem_ld_32x32b_x32(tmem_addr + 32, raw1);
    consume(raw0, 0);
    tmem_ld_wait();
    tmem_ld_32x32b_x32(tmem_addr + 64, raw0);
    consume(raw1, 1);
    tmem_ld_wait();
    tmem_ld_32x32b_x32(tmem_addr + 96, raw1);
    consume(raw0, 2);
    tmem_ld_wait();
    consume(raw1, 3);
#pragma unroll
    for (int i = 0; i < TT; ++i) out_h[(size_t)i * p.H] = acc_e[i];
#pragma unroll
    for (int i = 0; i < TU; ++i) out_h[(size_t)(TT + i) * p.H] = acc_d[i];
}

template <int TT_LOG2>
__global__ void __launch_bounds__(kDjThreads, 1)
dj_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_dy, const BwdParams p) {
    constexpr int TT = 1 << TT_LOG2, TU = 128 >> TT_LOG2;
    extern __shared__ __align__(1024) uint8_t smem[];
    const DjSmem L = dj_smem_layout();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* full = bars;              // [kDjStages]  leader: bytes of BOTH CTAs' loads
    uint64_t* empty = bars + 5;         // [kDjStages]  every CTA: the pair's MMAs are done with the stage
    uint64_t* acc_full = bars + 10;     // [2]          every CTA: accumulator buffer complete
    uint64_t* acc_empty = bars + 12;    // [2]          leader: 16 epilogue warps released the buffer
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L.tmem_off);
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster = (int)cluster_id_x(), n_clusters = (int)num_clusters_x();

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int i = 0; i < kDjStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
        fence_barrier_init();
    }
    if (warp_idx == kDjWarpLoad && lane == 0) { tma_prefetch_desc(&tmap_w); tma_prefetch_desc(&tmap_dy); }
    if (warp_idx == kDjWarpMma) tmem_alloc_2cta<512>(tmem_ptr);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

    const int n_tiles = p.tile_end - p.tile_begin;
    const int n_pairs = (n_tiles + 1) >> 1;
    const int n_units = n_pairs * p.NHC;  // consecutive units = the h-chunks of one tile pair (they share the dY images in L2)
    auto live_local = [&](int tl) { return tl < n_tiles && tile_live(p, p.tile_begin + tl); };

    if (warp_idx == kDjWarpLoad) {
        // ===================== loads (whole warp, converged; one lane elected per instruction) =====================
        uint32_t stage = 0, phase = 0;
        const uint32_t full0 = mapa_u32(smem_u32(&full[0]), 0);  // the leader's barriers
        for (int unit = cluster; unit < n_units; unit += n_clusters) {
            const int pi = unit / p.NHC, c = unit - pi * p.NHC;
            const bool live0 = live_local(2 * pi), live1 = live_local(2 * pi + 1);
            if (!live0 && !live1) continue;
            const int hbase = c * kDjChunkH;
            const int nbox0 = max(0, min(2, (p.H - hbase) >> 6)), nbox1 = max(0, min(2, (p.H - hbase - 128) >> 6));
            const bool my_live = rank ? live1 : live0;
            const int my_nbox = rank ? nbox1 : nbox0, h0 = hbase + rank * 128;
            const uint32_t bytes_pair = (uint32_t)(nbox0 + nbox1) * 8192u + (uint32_t)((int)live0 + (int)live1) * kImgBytes;
            const int img_row0 = (2 * pi + rank) * p.NT4 * 128;
            for (int kb = 0; kb < p.NVB; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1, 0x700 | stage);
                __syncwarp();
                uint8_t* st = smem + L.stage_off + stage * kDjStageBytes;
                if (rank == 0) mbar_arrive_expect_tx_e(&full[stage], bytes_pair);
                for (int j = 0; j < my_nbox; ++j) tma_load_2d_2cta_e(st + j * 8192, &tmap_w, full0 + stage * 8, h0 + j * 64, kb * 64);
                if (my_live) tma_load_2d_2cta_e(st + 2 * 8192, &tmap_dy, full0 + stage * 8, 0, img_row0 + kb * 128);
                if (++stage == kDjStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == kDjWarpMma) {
        // ===================== MMA issue (leader CTA; whole warp, converged) =====================
        if (rank == 0) {
            uint32_t stage = 0, phase = 0, it = 0;
            const uint32_t idesc = make_idesc_bf16(256, 256, 1, 0);
            const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem + L.stage_off), 8192, 1024);          // MN-major
            const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem + L.stage_off) + 2 * 8192, 0, 1024);   // K-major
            long long t_acc = 0, t_full = 0, tm = 0;
            const long long t_begin = clock64();
            for (int unit = cluster; unit < n_units; unit += n_clusters) {
                const int pi = unit / p.NHC;
                if (!live_local(2 * pi) && !live_local(2 * pi + 1)) continue;
                const uint32_t buf = it & 1;
                if (p.prof) tm = clock64();
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1, 0x800 | buf);
                if (p.prof) t_acc += clock64() - tm;
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int kb = 0; kb < p.NVB; ++kb) {
                    if (p.prof) tm = clock64();
                    mbar_wait(&full[stage], phase, 0x900 | stage);
                    if (p.prof) t_full += clock64() - tm;
                    tcgen05_fence_after();
                    const uint64_t off = (uint64_t)(stage * (kDjStageBytes >> 4));
                    umma_bf16_2cta_x4_mnA_e(d_tmem, a_desc0 + off, b_desc0 + off, idesc, kb != 0);
                    umma_commit_2cta_e(&empty[stage], 3);
                    if (++stage == kDjStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta_e(&acc_full[buf], 3);
                ++it;
            }
            if (p.prof && lane == 0) {
                long long* o = p.prof + blockIdx.x * 4;
                o[0] = clock64() - t_begin; o[1] = t_acc; o[2] = t_full; o[3] = it;
            }
        }
    } else {
        // ===================== epilogue: warps 0-7, lane quarter q, cell tile g of the pair =====================
        const int q = warp_idx & 3, g = warp_idx >> 2;
        const uint32_t acc_empty0 = mapa_u32(smem_u32(&acc_empty[0]), 0);
        const int per_b = p.nTt * p.nTu;
        uint32_t it = 0;
        for (int unit = cluster; unit < n_units; unit += n_clusters) {
            const int pi = unit / p.NHC, c = unit - pi * p.NHC;
            const bool live0 = live_local(2 * pi), live1 = live_local(2 * pi + 1);
            if (!live0 && !live1) continue;
            const uint32_t buf = it & 1;
            const int tl = 2 * pi + g;
            const int h = c * kDjChunkH + rank * 128 + q * 32 + lane;
            const bool work = (g ? live1 : live0) && (h - lane) < p.H;  // warp-uniform
            float e[TT], d[TU];
            if (work) {
                // the 24 pre-activation scalars of this thread's h (clamped rows: cells beyond T/U carry zero dY)
                const int tile = p.tile_begin + tl;
                const int b = tile / per_b, rem = tile - b * per_b;
                const int tt = rem / p.nTu, tu = rem - tt * p.nTu;
                const __nv_bfloat16* er = p.enc + ((size_t)b * p.T) * p.H + h;
                const __nv_bfloat16* dr = p.dec + ((size_t)b * p.U) * p.H + h;
#pragma unroll
                for (int i = 0; i < TT; ++i) e[i] = __bfloat162float(er[(size_t)min(tt * TT + i, p.T - 1) * p.H]);
#pragma unroll
                for (int i = 0; i < TU; ++i) d[i] = __bfloat162float(dr[(size_t)min(tu * TU + i, p.U - 1) * p.H]);
            }
            mbar_wait(&acc_full[buf], (it >> 1) & 1, 0xA00 | buf);
            tcgen05_fence_after();
            if (work) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + g * 128;
                float* out_h = p.dpre_part + (size_t)tl * (TT + TU) * p.H + h;
                switch (p.act_kind) {
                    case ACT_LEAKY_RELU: dj_epilogue_tile<TT_LOG2, ACT_LEAKY_RELU>(p, taddr, e, d, out_h); break;
                    case ACT_RELU: dj_epilogue_tile<TT_LOG2, ACT_RELU>(p, taddr, e, d, out_h); break;
                    case ACT_TANH: dj_epilogue_tile<TT_LOG2, ACT_TANH>(p, taddr, e, d, out_h); break;
                    default: dj_epilogue_tile<TT_LOG2, ACT_IDENTITY>(p, taddr, e, d, out_h); break;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank) mbar_arrive_cluster(acc_empty0 + buf * 8);
                else mbar_arrive(&acc_empty[buf]);
            }
            ++it;
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();  // the partner may still signal this CTA's barriers / read its shared memory
    if (warp_idx == kDjWarpMma) {
        tcgen05_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

// Fold the per-tile partial rows (flat, h-contiguous, one thread per output element).  d_enc rows are
// complete inside one (b, tt) group of label tiles (chunks are aligned to such groups) -> plain
// stores.  d_dec rows accumulate over tt inside the chunk and, by read-modify-write, across chunks.
__global__ void __launch_bounds__(256)
reduce_dpre_enc_kernel(const BwdParams p, float* __restrict__ d_enc) {
    const int tT = 1 << p.tT_log2, tU = kTileM >> p.tT_log2;
    const int g_begin = p.tile_begin / p.nTu, n_groups = (p.tile_end - p.tile_begin) / p.nTu;
    const long long total = (long long)n_groups * tT * p.H;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int h = (int)(o % p.H);
        const int ti = (int)((o / p.H) % tT);
        const int gl = (int)(o / ((long long)p.H * tT));
        const int g = g_begin + gl;
        const int b = g / p.nTt, tt = g - b * p.nTt;
        const int t = tt * tT + ti;
        if (t >= p.logit_lengths[b]) continue;  // rows beyond T_b stay zero (memset)
        const int n_tu = (p.target_lengths[b] + 1 + tU - 1) / tU;  // live label tiles
        const float* base = p.dpre_part + ((size_t)gl * p.nTu * (tT + tU) + ti) * p.H + h;
        float s = 0.f;
        for (int tu = 0; tu < n_tu; ++tu) s += base[(size_t)tu * (tT + tU) * p.H];
        d_enc[((size_t)b * p.T + t) * p.H + h] = s;
    }
}

__global__ void __launch_bounds__(256)
reduce_dpre_dec_kernel(const BwdParams p, float* __restrict__ d_dec) {
    const int tT = 1 << p.tT_log2, tU = kTileM >> p.tT_log2;
    const int g_begin = p.tile_begin / p.nTu, g_end = p.tile_end / p.nTu;
    const int b_begin = g_begin / p.nTt, b_end = (g_end - 1) / p.nTt + 1;
    const long long total = (long long)(b_end - b_begin) * p.U * p.H;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int h = (int)(o % p.H);
        const int u = (int)((o / p.H) % p.U);
        const int b = b_begin + (int)(o / ((long long)p.H * p.U));
        const int Tb = p.logit_lengths[b];
        if (u >= p.target_lengths[b] + 1) continue;
        const int tu = u / tU, ui = u - tu * tU;
        const int gb0 = max(g_begin, b * p.nTt), gb1 = min(g_end, (b + 1) * p.nTt);
        float s = 0.f;
        for (int g = gb0; g < gb1; ++g) {
            if ((g - b * p.nTt) * tT >= Tb) break;  // dead frame tiles carry no data
            s += p.dpre_part[((size_t)(g * p.nTu + tu - p.tile_begin) * (tT + tU) + tT + ui) * p.H + h];
        }
        d_dec[((size_t)b * p.U + u) * p.H + h] += s;
    }
}

// ------------------------------------------------------------------------------------------------
// dW / db GEMM (contraction over cells), split-K over the chunk's tiles
// ------------------------------------------------------------------------------------------------
static constexpr int kDwStages = 2;
static constexpr int kDwStageBytes = 6 * kImgBytes;  // 2 dY v-block images + up to 4 J h-block images

struct DwSmem { uint32_t stage_off, ones_off, bar_off, tmem_off, total; };
__host__ __device__ inline DwSmem dw_smem_layout() {
    DwSmem l;
    l.stage_off = 0;
    l.ones_off = kDwStages * kDwStageBytes;
    l.bar_off = l.ones_off + kImgBytes;
    l.tmem_off = l.bar_off + 8 * 8;
    l.total = l.tmem_off + 16;
    return l;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
dw_gemm_kernel(const BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const DwSmem L = dw_smem_layout();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* full = bars;        // [2]
    uint64_t* empty = bars + 2;   // [2]
    uint64_t* acc_full = bars + 4;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L.tmem_off);
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // unit = (v-tile, h-tile, split)
    const int unit = blockIdx.x;
    const int split = unit % p.n_splits;
    const int ht = (unit / p.n_splits) % p.NHT;
    const int vt = unit / (p.n_splits * p.NHT);
    const int hb0 = ht * 4, nhb = min(4, p.KB - hb0);
    const int nvb = min(2, p.NVB - vt * 2);          // v-block images that exist for this v-tile
    const bool with_db = ht == p.NHT - 1;            // the last h-tile also carries the bias gradient
    const int n_tiles = p.tile_end - p.tile_begin;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int i = 0; i < kDwStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    // constant "ones" B block: J-image layout with column h = 0 set to 1.0 for all 128 cells
    {
        uint4* o = reinterpret_cast<uint4*>(smem + L.ones_off);
        for (int i = threadIdx.x; i < kImgBytes / 16; i += blockDim.x) o[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        if (threadIdx.x < 128)
            *reinterpret_cast<uint16_t*>(smem + L.ones_off + sw128_offset(threadIdx.x, 0)) = 0x3F80;  // bf16 1.0
        fence_proxy_async_smem();
    }
    if (warp_idx == kBwdWarpMma) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

    if (warp_idx == kBwdWarpLoad) {
        {
            uint32_t stage = 0, phase = 0;
            for (int tl = split; tl < n_tiles; tl += p.n_splits) {
                if (!tile_live(p, p.tile_begin + tl)) continue;
                mbar_wait(&empty[stage], phase ^ 1, 0xB00 | stage);
                __syncwarp();
                uint8_t* st = smem + L.stage_off + stage * kDwStageBytes;
                const uint8_t* dy = reinterpret_cast<const uint8_t*>(p.dY_img) + ((size_t)tl * p.NT4 + vt * 2) * kImgBytes;
                const uint8_t* jm = reinterpret_cast<const uint8_t*>(p.J_img) + ((size_t)tl * p.KB + hb0) * kImgBytes;
                mbar_arrive_expect_tx_e(&full[stage], (2 + nhb) * kImgBytes);
                bulk_load_1d_e(st, dy, kImgBytes, &full[stage]);
                // a v-tile whose second 64-column block lies beyond V re-reads the first block: those
                // accumulator rows (v >= V) are never stored
                bulk_load_1d_e(st + kImgBytes, dy + (nvb == 2 ? kImgBytes : 0), kImgBytes, &full[stage]);
                bulk_load_1d_e(st + 2 * kImgBytes, jm, nhb * kImgBytes, &full[stage]);
                if (++stage == kDwStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == kBwdWarpMma) {
        {  // whole warp, converged
            uint32_t stage = 0, phase = 0;
            bool first = true;
            const uint32_t idesc = make_idesc_bf16(kTileM, nhb * 64, 1, 1);
            const uint32_t idesc_db = make_idesc_bf16(kTileM, 16, 1, 1);
            const uint32_t ones_base = smem_u32(smem + L.ones_off);
            for (int tl = split; tl < n_tiles; tl += p.n_splits) {
                if (!tile_live(p, p.tile_begin + tl)) continue;
                mbar_wait(&full[stage], phase, 0xC00 | stage);
                tcgen05_fence_after();
                const uint32_t a_base = smem_u32(smem + L.stage_off + stage * kDwStageBytes);
                const uint32_t b_base = a_base + 2 * kImgBytes;
#pragma unroll
                for (int k = 0; k < 8; ++k) {  // 16 cells per step = two 8-row groups of the image
                    const uint64_t a_desc = make_smem_desc_sw128(a_base + k * 2048, kImgBytes, 1024);  // MN-major, M = v
                    const uint64_t b_desc = make_smem_desc_sw128(b_base + k * 2048, kImgBytes, 1024);  // MN-major, N = h
                    umma_bf16_e(tmem_base, a_desc, b_desc, idesc, !(first && k == 0));
                    if (with_db) {
                        const uint64_t o_desc = make_smem_desc_sw128(ones_base + k * 2048, kImgBytes, 1024);
                        umma_bf16_e(tmem_base + 256, a_desc, o_desc, idesc_db, !(first && k == 0));
                    }
                }
                first = false;
                umma_commit_e(&empty[stage]);
                if (++stage == kDwStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_e(acc_full);
        }
    } else {
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        const int v = vt * 128 + row;
        // did this split see any live tile?  (all roles agree; the epilogue must not wait otherwise)
        bool any = false;
        for (int tl = split; tl < n_tiles; tl += p.n_splits) any |= tile_live(p, p.tile_begin + tl);
        const size_t vrows = (size_t)p.NVT * 128;
        float* wrow = p.dW_part + ((size_t)split * vrows + v) * p.H + hb0 * 64;
        if (any) {
            mbar_wait(acc_full, 0, 0xD00);
            tcgen05_fence_after();
        }
        for (int cc = 0; cc < nhb * 64; cc += 32) {
            uint32_t raw[32];
            if (any) {
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + cc, raw);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) raw[j] = 0u;
            }
            float4* dst = reinterpret_cast<float4*>(wrow + cc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 x = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                       __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
                if (p.accumulate) {
                    const float4 o = dst[j];
                    x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w;
                }
                dst[j] = x;
            }
        }
        if (with_db) {
            uint32_t raw[32];
            float x = 0.f;
            if (any) {
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + 256, raw);
                tmem_ld_wait();
                x = __uint_as_float(raw[0]);
            }
            float* d = p.db_part + (size_t)split * vrows + v;
            *d = p.accumulate ? *d + x : x;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == kBwdWarpMma) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

__global__ void reduce_dw_kernel(const BwdParams p, float* __restrict__ dW, float* __restrict__ db) {
    const size_t vrows = (size_t)p.NVT * 128;
    const size_t n = (size_t)p.V * p.H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < p.n_splits; ++sp) s += p.dW_part[(size_t)sp * vrows * p.H + i];
        dW[i] = s;
    }
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < (size_t)p.V; v += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < p.n_splits; ++sp) s += p.db_part[(size_t)sp * vrows + v];
        db[v] = s;
    }
}

}  // namespace tsasr
