// backward_gemm.cuh -- the two backward GEMMs of the fused joint (tcgen05 + TMEM + TMA/bulk copies).
//
// They consume the operand images the MODE_GRAD pass of joint_gemm.cuh leaves in the chunk workspace: per
// 128-cell tile, dY images [v-block][128 cells x 64 v] and J images [h-block][128 cells x 64 h], both bf16 in
// the SWIZZLE_128B shared-memory image, so an un-swizzled TMA box copies image rows verbatim and the block is
// directly addressable by a UMMA descriptor -- K-major when the 64-wide dimension is the contraction (dJ),
// MN-major when the 128 cells are the contraction (dW).  Both kernels run on CTA pairs (cta_group::2) and walk
// the live -- or, with tile pruning, the active -- tiles only (LiveCursor, common.cuh).
//
//   dj_gemm_kernel : dJ^T[h, cells] = sum_v W[v, h] * dY[cells, v]          (autograd of SB/nnet/linear.py:74)
//                    epilogue: dpre = dJ * act'(enc + dec) (transducer_joint.py:95 backward) and the two
//                    broadcast-sum reductions of transducer_joint.py:74 (sum over u -> d_enc rows,
//                    sum over t -> d_dec rows) as in-register adds; per-tile partial rows go to a small
//                    buffer that reduce_dpre_kernel folds deterministically.
//   dw_gemm_kernel : dW[v, h] += sum_cells dY[cells, v] * J[cells, h], db[v] += sum_cells dY[cells, v]
//                    (db rides along as 32 extra accumulator columns fed by a constant "ones" block).
//   tile_activity_kernel / compact_active_kernel : backward tile pruning.
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
    float* dW_part;    // [max splits][NV2 * 256][H]
    float* db_part;    // [max splits][NV2 * 256]
    int NV2;           // ceil(V / 256): v-tiles of a CTA pair
    int NHU;           // h-units (1 or 2); the last one also carries db
    int hu_blk[3];     // unit u covers h-blocks [hu_blk[u], hu_blk[u+1])
    int hu_splits[2];  // split-K factor of unit u
    int hu_item0[3];   // work items of h-unit u are numbered [hu_item0[u], hu_item0[u+1])
    int accumulate;    // 0: store, 1: read-modify-write (later chunks)
    long long* prof;   // development: MMA-warp cycle counters per CTA (or nullptr)
    // backward tile pruning (nullptr / nullptr: every live tile is processed)
    const int* active_ids;       // ordered dense ids of the chunk's active tiles
    const int* active_count;     // their number
    const uint8_t* tile_flags;   // [tile - tile_begin] 1 = active (what the fold of the partial rows may read)
};

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
// act'(x) of the pre-activation x = enc + dec, as the reference's autograd evaluates it (fp32, from the UNROUNDED
// activation: torch's LeakyReLU / ReLU backward test x > 0, tanh backward is 1 - tanh(x)^2); the bf16 rounding of the
// GEMM operand J is a straight-through step and has no derivative of its own.
template <int ACT>
__device__ __forceinline__ float act_grad_pre(float x, float param) {
    if (ACT == ACT_LEAKY_RELU) return x > 0.f ? 1.f : param;
    if (ACT == ACT_RELU) return x > 0.f ? 1.f : 0.f;
    if (ACT == ACT_TANH) {
        const float j = tanhf(x);
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
    griddep_wait();  // programmatic dependent launch: the dY images come from the preceding GRAD pass
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
    // unit = (pair of consecutive LIVE tiles, h-chunk); consecutive units = the h-chunks of one tile pair (they share
    // the dY images in L2).  tl0 / tl1: chunk-local ids of the pair's tiles (tl1 = -1: the last pair is incomplete).
    const int n_live = count_live_tiles_warp(p);  // every warp counts for itself (converged here)
    LiveCursor<BwdParams> cur(p, n_live);
    auto open_unit = [&](int unit, int& c, int& tl0, int& tl1) -> bool {
        const int pi = unit / p.NHC;
        c = unit - pi * p.NHC;
        if (!cur.seek(p, 2 * pi)) return false;
        tl0 = cur.tile(p) - p.tile_begin;
        tl1 = cur.seek(p, 2 * pi + 1) ? cur.tile(p) - p.tile_begin : -1;
        cur.hint_pair(p, 2 * ((unit + n_clusters) / p.NHC));  // list mode: the next unit's ids are fetched early
        return true;
    };

    if (warp_idx == kDjWarpLoad) {
        // ===================== loads (whole warp, converged; one lane elected per instruction) =====================
        uint32_t stage = 0, phase = 0;
        const uint32_t full0 = mapa_u32(smem_u32(&full[0]), 0);  // the leader's barriers
        for (int unit = cluster;; unit += n_clusters) {
            int c, tl0, tl1;
            if (!open_unit(unit, c, tl0, tl1)) break;
            const int hbase = c * kDjChunkH;
            const int nbox0 = max(0, min(2, (p.H - hbase) >> 6)), nbox1 = max(0, min(2, (p.H - hbase - 128) >> 6));
            const int my_tl = rank ? tl1 : tl0;
            const bool my_live = my_tl >= 0;
            const int my_nbox = rank ? nbox1 : nbox0, h0 = hbase + rank * 128;
            const uint32_t bytes_pair = (uint32_t)(nbox0 + nbox1) * 8192u + (uint32_t)(1 + (tl1 >= 0)) * kImgBytes;
            const int img_row0 = my_tl * p.NT4 * 128;
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
            // the issuing warp needs no tile coordinates, only the number of units: count the live tiles once
            const int n_units = ((n_live + 1) >> 1) * p.NHC;
            for (int unit = cluster; unit < n_units; unit += n_clusters) {
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
        for (int unit = cluster;; unit += n_clusters) {
            int c, tl0, tl1;
            if (!open_unit(unit, c, tl0, tl1)) break;
            const uint32_t buf = it & 1;
            const int tl = g ? tl1 : tl0;
            const int h = c * kDjChunkH + rank * 128 + q * 32 + lane;
            const bool work = tl >= 0 && (h - lane) < p.H;  // warp-uniform
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

// ------------------------------------------------------------------------------------------------
// Backward tile pruning.  Every term of d cost / d logits of cell (t,u) carries the factor
// exp(alpha(t,u) + beta(t,u) - L) (or less: the blank / emit terms are parts of it), the posterior probability that
// the alignment passes through the cell.  Away from the band of plausible alignments it decays like a Gaussian
// tail (1e-100 and below at config 2), so whole tiles contribute nothing that survives fp32 accumulation next to
// the O(1) terms of the band -- let alone the bf16 rounding of dlogits.  A tile is ACTIVE when its largest
// relative occupancy is >= 2^prune_log2_eps (default 2^-30); only active tiles are recomputed and fed to the
// backward GEMMs.  NaN occupancies (non-finite logits) keep their tile active, so they still poison the
// utterance's gradients as in the reference.
// ------------------------------------------------------------------------------------------------
// one warp per tile of the chunk: flags[tile - tile_begin] = 1 (active) / 0
__global__ void __launch_bounds__(256)
tile_activity_kernel(const BwdParams p, const float* __restrict__ alpha, const float* __restrict__ beta,
                     const float* __restrict__ cost, float ln_eps, uint8_t* __restrict__ flags) {
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int tl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tl >= p.tile_end - p.tile_begin) return;
    const int tile = p.tile_begin + tl;
    const int per_b = p.nTt * p.nTu;
    const int b = tile / per_b, rem = tile - b * per_b;
    const int tt = rem / p.nTu, tu = rem - tt * p.nTu;
    const int tT = 1 << p.tT_log2, tU = kTileM >> p.tT_log2;
    const int Tb = min(max(p.logit_lengths[b], 1), p.T), Ub = min(max(p.target_lengths[b], 0), p.U - 1) + 1;
    bool active = false;
    if (tt * tT < Tb && tu * tU < Ub) {  // a live tile
        const float L = -cost[b];
        float m = -INFINITY;
        bool nan = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            const int t = tt * tT + (r & (tT - 1)), u = tu * tU + (r >> p.tT_log2);
            if (t < Tb && u < Ub) {
                const size_t o = skew_index(b, t, u, p.T, p.U);
                const float x = alpha[o] + beta[o] - L;
                nan |= !(x == x);
                m = fmaxf(m, x);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        nan = __any_sync(0xffffffffu, nan);
        active = nan || !(m < ln_eps);
    }
    if (lane == 0) flags[tl] = active ? 1 : 0;
}

// ordered compaction of the active tiles (single CTA): ids[0..count) = dense tile ids, stats[0] = count of this chunk,
// stats[1] += count, stats[2] += live tiles (both cumulative over the chunks of one backward call)
__global__ void __launch_bounds__(1024)
compact_active_kernel(const BwdParams p, const uint8_t* __restrict__ flags, int* __restrict__ ids, int* __restrict__ stats) {
    __shared__ int warp_sums[32];
    __shared__ int running;
    griddep_wait();
    const int n = p.tile_end - p.tile_begin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const bool f = i < n && flags[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_sums[warp] = __popc(bal);
        __syncthreads();
        int off = running;
        for (int w = 0; w < warp; ++w) off += warp_sums[w];
        if (f) ids[off + __popc(bal & ((1u << lane) - 1u))] = p.tile_begin + i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += warp_sums[w];
            running += t;
        }
        __syncthreads();
    }
    if (warp == 0) {
        const int live = count_live_tiles_warp_geometric(p);
        if (lane == 0) { stats[0] = running; stats[1] += running; stats[2] += live; }
    }
}

// Fold the per-tile partial rows: one block row per output row, one thread per 4 consecutive h (float4).
// blockIdx.x < n_enc_rows: d_enc row (frame-tile group gl, ti) = sum over the live label tiles of the group --
// rows are complete inside a chunk (chunks are aligned to whole (utterance, frame-tile) groups) -> plain stores.
// Other blocks: d_dec row (b, u) = sum over the chunk's frame tiles of utterance b, accumulated across
// chunks by read-modify-write (d_dec is zero-filled first).  Deterministic, no atomics.
__global__ void __launch_bounds__(160)
reduce_dpre_kernel(const BwdParams p, float* __restrict__ d_enc, float* __restrict__ d_dec) {
    griddep_wait();
    const int tT = 1 << p.tT_log2, tU = kTileM >> p.tT_log2;
    const int g_begin = p.tile_begin / p.nTu, g_end = p.tile_end / p.nTu;
    const int n_enc_rows = (g_end - g_begin) * tT;
    const int H4 = p.H >> 2;
    const size_t tile_stride4 = (size_t)(tT + tU) * H4;  // float4 units between consecutive tiles
    const float4* part = reinterpret_cast<const float4*>(p.dpre_part);
    if ((int)blockIdx.x < n_enc_rows) {
        const int gl = blockIdx.x >> p.tT_log2, ti = blockIdx.x & (tT - 1);
        const int g = g_begin + gl;
        const int b = g / p.nTt, tt = g - b * p.nTt;
        const int t = tt * tT + ti;
        if (t >= min(max(p.logit_lengths[b], 1), p.T)) return;  // rows beyond T_b stay zero (memset)
        const int n_tu = (min(max(p.target_lengths[b], 0), p.U - 1) + 1 + tU - 1) / tU;  // live label tiles
        for (int h4 = threadIdx.x; h4 < H4; h4 += blockDim.x) {
            const float4* base = part + ((size_t)gl * p.nTu) * tile_stride4 + (size_t)ti * H4 + h4;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int tu = 0; tu < n_tu; ++tu) {
                if (p.tile_flags && !p.tile_flags[gl * p.nTu + tu]) continue;  // pruned tile: no partial rows exist
                const float4 x = base[(size_t)tu * tile_stride4];
                s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
            }
            reinterpret_cast<float4*>(d_enc)[((size_t)b * p.T + t) * H4 + h4] = s;
        }
    } else {
        const int row = blockIdx.x - n_enc_rows;
        const int b_begin = g_begin / p.nTt;
        const int b = b_begin + row / p.U, u = row % p.U;
        if (u >= min(max(p.target_lengths[b], 0), p.U - 1) + 1) return;
        const int Tb = min(max(p.logit_lengths[b], 1), p.T);
        const int tu = u / tU, ui = u - tu * tU;
        const int gb0 = max(g_begin, b * p.nTt), gb1 = min(min(g_end, (b + 1) * p.nTt), b * p.nTt + (Tb + tT - 1) / tT);
        for (int h4 = threadIdx.x; h4 < H4; h4 += blockDim.x) {
            const float4* base = part + ((size_t)(gb0 * p.nTu + tu - p.tile_begin)) * tile_stride4 + (size_t)(tT + ui) * H4 + h4;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int g = gb0; g < gb1; ++g) {  // dead frame tiles carry no data
                if (p.tile_flags && !p.tile_flags[g * p.nTu + tu - p.tile_begin]) continue;  // pruned tile
                const float4 x = base[(size_t)(g - gb0) * p.nTu * tile_stride4];
                s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
            }
            float4* o = reinterpret_cast<float4*>(d_dec) + ((size_t)b * p.U + u) * H4 + h4;
            float4 y = *o;
            y.x += s.x; y.y += s.y; y.z += s.z; y.w += s.w;
            *o = y;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dW / db GEMM (contraction over cells), split-K over the chunk's tiles, on CTA pairs
// ------------------------------------------------------------------------------------------------
// dW[v, h] = sum_cells dY[cell, v] * J[cell, h] as a cta_group::2 MMA with M = 256 v-rows (each CTA owns 128 =
// two dY images, MN-major A) and N = up to 512 h-columns (the pair's whole TMEM width; each CTA supplies half
// of the J images, MN-major B).  The accumulator lives in TMEM for the whole kernel (split-K over the tiles of
// the chunk); a unit is (256 v-rows, h-unit, split).  H is cut into at most two h-units so that the last one
// leaves room for db: 32 extra accumulator columns fed by a constant "ones" block (db[v] = sum_cells dY[cell, v]).
// A pipeline stage is HALF a cell tile (64 cells: 8 KB half-images), four stages of fixed geometry (narrow units
// leave part of each stage unused; a deeper ring for them was measured and makes no difference).
static constexpr int kDwStages = 4;
static constexpr int kDwHalfImg = kImgBytes / 2;          // rows 0-63 or 64-127 of an operand image
static constexpr int kDwStageBytes = 6 * kDwHalfImg;      // 2 dY + up to 4 J half-images
static constexpr int kDwRingBytes = kDwStages * kDwStageBytes;  // 192 KB
static constexpr int kDwDbCols = 32;

struct DwSmem { uint32_t stage_off, ones_off, bar_off, tmem_off, total; };
__host__ __device__ inline DwSmem dw_smem_layout() {
    DwSmem l;
    l.stage_off = 0;
    l.ones_off = kDwRingBytes;
    l.bar_off = l.ones_off + kDwHalfImg;
    l.tmem_off = l.bar_off + 24 * 8;
    l.total = l.tmem_off + 16;
    return l;
}

// One pipeline stage of the dW kernel in ONE asm statement: four K = 16 steps (all descriptors advance by 16
// rows = 2048 B), each with up to three MMAs sharing the A operand: dW columns [0,256) (idesc0), dW columns
// [256,512) from the J blocks two half-images further (idesc1, if has1) and the db columns fed by the constant
// ones block (idesc_db, if has_db).  Issued by a converged warp, one lane elected once.
__device__ __forceinline__ void umma_dw_stage_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint64_t o_desc, uint32_t idesc0,
                                                uint32_t idesc1, uint32_t idesc_db, uint32_t accumulate_first, uint32_t has1,
                                                uint32_t has_db, uint32_t db_col) {
    asm volatile(
        "{\n\t.reg .pred p, q, t, h1, hd;\n\t"
        ".reg .b64 a1, a2, a3, b1, b2, b3, c0, c1, c2, c3, o1, o2, o3;\n\t.reg .b32 d1, dd;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "setp.ne.b32 h1, %8, 0;\n\tand.pred h1, h1, q;\n\t"
        "setp.ne.b32 hd, %9, 0;\n\tand.pred hd, hd, q;\n\t"
        "add.u32 d1, %0, 256;\n\tadd.u32 dd, %0, %10;\n\t"
        "add.u64 a1, %1, 128;\n\tadd.u64 a2, %1, 256;\n\tadd.u64 a3, %1, 384;\n\t"
        "add.u64 b1, %2, 128;\n\tadd.u64 b2, %2, 256;\n\tadd.u64 b3, %2, 384;\n\t"
        "add.u64 c0, %2, 1024;\n\tadd.u64 c1, %2, 1152;\n\tadd.u64 c2, %2, 1280;\n\tadd.u64 c3, %2, 1408;\n\t"
        "add.u64 o1, %3, 128;\n\tadd.u64 o2, %3, 256;\n\tadd.u64 o3, %3, 384;\n\t"
        "@q  tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %4, p;\n\t"
        "@h1 tcgen05.mma.cta_group::2.kind::f16 [d1], %1, c0, %5, p;\n\t"
        "@hd tcgen05.mma.cta_group::2.kind::f16 [dd], %1, %3, %6, p;\n\t"
        "@q  tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %4, t;\n\t"
        "@h1 tcgen05.mma.cta_group::2.kind::f16 [d1], a1, c1, %5, t;\n\t"
        "@hd tcgen05.mma.cta_group::2.kind::f16 [dd], a1, o1, %6, t;\n\t"
        "@q  tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %4, t;\n\t"
        "@h1 tcgen05.mma.cta_group::2.kind::f16 [d1], a2, c2, %5, t;\n\t"
        "@hd tcgen05.mma.cta_group::2.kind::f16 [dd], a2, o2, %6, t;\n\t"
        "@q  tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %4, t;\n\t"
        "@h1 tcgen05.mma.cta_group::2.kind::f16 [d1], a3, c3, %5, t;\n\t"
        "@hd tcgen05.mma.cta_group::2.kind::f16 [dd], a3, o3, %6, t;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "l"(o_desc), "r"(idesc0), "r"(idesc1), "r"(idesc_db), "r"(accumulate_first),
          "r"(has1), "r"(has_db), "r"(db_col) : "memory");
}

// Work item = (h-unit, v-tile of 256 rows, split-K slice); items are numbered h-unit 0 first.  CTA pair c works
// on items c, c + #pairs, ... : with one item per pair this is a plain split-K grid (config 2); with several
// (large V) the host picks the split factors so that every pair's items add up to about the same cost.
struct DwItem { int hu, vt2, split, n_splits, hb0, nblk, nblk_e, nb0, nb1, n_slots; bool with_db; };
__device__ __forceinline__ DwItem dw_item(const BwdParams& p, int item) {
    DwItem it;
    it.hu = (p.NHU > 1 && item >= p.hu_item0[1]) ? 1 : 0;
    const int local = item - p.hu_item0[it.hu];
    it.n_splits = p.hu_splits[it.hu];
    it.split = local % it.n_splits;
    it.vt2 = local / it.n_splits;
    it.hb0 = p.hu_blk[it.hu];
    it.nblk = p.hu_blk[it.hu + 1] - it.hb0;     // real h-blocks of the unit
    it.nblk_e = (it.nblk + 1) & ~1;             // even: the pair's CTAs supply equal halves
    it.nb0 = min(4, it.nblk_e);                 // blocks per MMA (N = 64 * nb)
    it.nb1 = it.nblk_e - it.nb0;
    it.n_slots = (it.nb0 >> 1) + (it.nb1 >> 1);  // J half-images per CTA and stage
    it.with_db = it.hu == p.NHU - 1;
    return it;
}
// J h-block held in shared slot j (= 2 * mma + i) of CTA r: D columns come out in natural h order
__device__ __forceinline__ int dw_slot_block(const BwdParams& p, const DwItem& it, int r, int j) {
    const int m = j >> 1, i = j & 1;
    const int nb = m ? it.nb1 : it.nb0;
    if (i >= (nb >> 1)) return -1;
    const int hb = it.hb0 + 4 * m + r * (nb >> 1) + i;
    return hb < p.KB ? hb : -1;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
dw_gemm_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_j, const BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_wait();  // programmatic dependent launch
    const DwSmem L = dw_smem_layout();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* full = bars;        // [kDwStages]  leader: bytes of both CTAs
    uint64_t* empty = bars + 8;   // [kDwStages]  every CTA
    uint64_t* acc_full = bars + 16;   // every CTA: the item's accumulator is complete
    uint64_t* acc_empty = bars + 17;  // leader: 8 epilogue warps have drained it
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L.tmem_off);
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int pair = (int)cluster_id_x(), n_pairs = (int)num_clusters_x();
    const int n_items = p.hu_item0[p.NHU];
    const int n_live = count_live_tiles_warp(p);  // every warp counts for itself (converged here)

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int i = 0; i < kDwStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        fence_barrier_init();
    }
    // constant "ones" B block (leader: column 0 = 1.0 for all 64 cells; partner: zeros)
    {
        uint4* o = reinterpret_cast<uint4*>(smem + L.ones_off);
        for (int i = threadIdx.x; i < kDwHalfImg / 16; i += blockDim.x) o[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        if (rank == 0 && threadIdx.x < 64)
            *reinterpret_cast<uint16_t*>(smem + L.ones_off + sw128_offset(threadIdx.x, 0)) = 0x3F80;  // bf16 1.0
        fence_proxy_async_smem();
    }
    if (warp_idx == kBwdWarpLoad && lane == 0) { tma_prefetch_desc(&tmap_dy); tma_prefetch_desc(&tmap_j); }
    if (warp_idx == kBwdWarpMma) tmem_alloc_2cta<512>(tmem_ptr);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    if (warp_idx == kBwdWarpLoad) {
        uint32_t stage = 0, phase = 0;
        const uint32_t full0 = mapa_u32(smem_u32(&full[0]), 0);
        for (int item = pair; item < n_items; item += n_pairs) {
            const DwItem it = dw_item(p, item);
            int nB = 0;
            for (int r = 0; r < 2; ++r)
                for (int j = 0; j < 4; ++j) nB += dw_slot_block(p, it, r, j) >= 0;
            const uint32_t bytes_pair = (uint32_t)(4 + nB) * kDwHalfImg;
            LiveCursor<BwdParams> cur(p, n_live);
            for (int k = it.split;; k += it.n_splits) {  // the item's share of the LIVE tiles
                if (!cur.seek(p, k)) break;
                cur.hint(p, k + it.n_splits);  // list mode: next id in flight while this tile's stages are issued
                const int tl = cur.tile(p) - p.tile_begin;
                const int dy_row0 = (tl * p.NT4 + it.vt2 * 4 + rank * 2) * 128;
                for (int half = 0; half < 2; ++half) {
                    mbar_wait(&empty[stage], phase ^ 1, 0xB00 | stage);
                    __syncwarp();
                    uint8_t* st = smem + L.stage_off + stage * kDwStageBytes;
                    if (rank == 0) mbar_arrive_expect_tx_e(&full[stage], bytes_pair);
                    tma_load_2d_2cta_e(st, &tmap_dy, full0 + stage * 8, 0, dy_row0 + half * 64);
                    tma_load_2d_2cta_e(st + kDwHalfImg, &tmap_dy, full0 + stage * 8, 0, dy_row0 + 128 + half * 64);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int hb = dw_slot_block(p, it, rank, j);
                        if (hb >= 0) tma_load_2d_2cta_e(st + (2 + j) * kDwHalfImg, &tmap_j, full0 + stage * 8, 0, (tl * p.KB + hb) * 128 + half * 64);
                    }
                    if (++stage == kDwStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == kBwdWarpMma) {
        if (rank == 0) {  // whole warp, converged
            uint32_t stage = 0, phase = 0, n_done = 0;
            const uint32_t idesc_db = make_idesc_bf16(256, kDwDbCols, 1, 1);
            // MN-major operands: 64-wide blocks kDwHalfImg apart, 8-row K groups 1024 B apart; a k-step is 16 rows
            const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem + L.stage_off), kDwHalfImg, 1024);
            const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem + L.stage_off) + 2 * kDwHalfImg, kDwHalfImg, 1024);
            const uint64_t o_desc0 = make_smem_desc_sw128(smem_u32(smem + L.ones_off), kDwHalfImg, 1024);
            long long t_full = 0, tm = 0, n_st = 0;
            const long long t_begin = clock64();
            for (int item = pair; item < n_items; item += n_pairs) {
                const DwItem it = dw_item(p, item);
                const uint32_t idesc0 = make_idesc_bf16(256, it.nb0 * 64, 1, 1);
                const uint32_t idesc1 = make_idesc_bf16(256, it.nb1 > 0 ? it.nb1 * 64 : 64, 1, 1);
                bool first = true;
                for (int k = it.split; k < n_live; k += it.n_splits) {  // the item's share of the live tiles
                    if (first) {  // the previous item's accumulator must have been drained by both CTAs
                        mbar_wait(acc_empty, (n_done & 1) ^ 1, 0xC80);
                        tcgen05_fence_after();
                    }
                    for (int half = 0; half < 2; ++half) {
                        if (p.prof) tm = clock64();
                        mbar_wait(&full[stage], phase, 0xC00 | stage);
                        if (p.prof) { t_full += clock64() - tm; ++n_st; }
                        tcgen05_fence_after();
                        // 4 k-steps x (dW columns 0-255, dW columns 256-511, db) in one issue block
                        const uint32_t off = stage * (kDwStageBytes >> 4);
                        umma_dw_stage_e(tmem_base, a_desc0 + off, b_desc0 + off, o_desc0, idesc0, idesc1, idesc_db, !first,
                                        it.nb1 > 0, it.with_db, it.nblk_e * 64);
                        first = false;
                        umma_commit_2cta_e(&empty[stage], 3);
                        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
                    }
                }
                if (!first) {  // at least one live tile: hand the accumulator to the epilogues
                    umma_commit_2cta_e(acc_full, 3);
                    ++n_done;
                }
            }
            if (p.prof && lane == 0) {
                long long* o = p.prof + blockIdx.x * 4;
                o[0] = clock64() - t_begin; o[1] = 0; o[2] = t_full; o[3] = n_st;
            }
        }
    } else {
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        const uint32_t acc_empty_leader = mapa_u32(smem_u32(acc_empty), 0);
        const size_t vrows = (size_t)p.NV2 * 256;
        uint32_t n_done = 0;
        for (int item = pair; item < n_items; item += n_pairs) {
            const DwItem it = dw_item(p, item);
            const int v = it.vt2 * 256 + rank * 128 + row;
            // does this split own any live tile?  (all roles agree; the epilogue must not wait otherwise)
            const bool any = it.split < n_live;
            float* wrow = p.dW_part + ((size_t)it.split * vrows + v) * p.H + it.hb0 * 64;
            if (any) {
                mbar_wait(acc_full, n_done & 1, 0xD00);
                tcgen05_fence_after();
            }
            for (int cc = 0; cc < it.nblk * 64; cc += 32) {
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
            if (it.with_db) {
                uint32_t raw[16];
                float x = 0.f;
                if (any) {
                    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + it.nblk_e * 64, raw);
                    tmem_ld_wait();
                    x = __uint_as_float(raw[0]);
                }
                float* d = p.db_part + (size_t)it.split * vrows + v;
                *d = p.accumulate ? *d + x : x;
            }
            if (any) {  // the accumulator may be overwritten by the pair's next item
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (rank) mbar_arrive_cluster(acc_empty_leader);
                    else mbar_arrive(acc_empty);
                }
                ++n_done;
            }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp_idx == kBwdWarpMma) {
        tcgen05_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

__global__ void __launch_bounds__(256)
reduce_dw_kernel(const BwdParams p, float* __restrict__ dW, float* __restrict__ db) {
    griddep_wait();
    const size_t vrows = (size_t)p.NV2 * 256;
    const size_t n4 = (size_t)p.V * p.H / 4;  // H % 64 == 0: a float4 never straddles rows or h-units
    const float4* part = reinterpret_cast<const float4*>(p.dW_part);
    const size_t split_stride4 = vrows * p.H / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int h = (int)((i * 4) % p.H);
        const int ns = p.hu_splits[(p.NHU > 1 && h >= p.hu_blk[1] * 64) ? 1 : 0];
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int sp = 0; sp < ns; ++sp) {
            const float4 x = part[(size_t)sp * split_stride4 + i];
            s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
        }
        reinterpret_cast<float4*>(dW)[i] = s;
    }
    const int ns_db = p.hu_splits[p.NHU - 1];
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < (size_t)p.V; v += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < ns_db; ++sp) s += p.db_part[(size_t)sp * vrows + v];
        db[v] = s;
    }
}

}  // namespace tsasr
