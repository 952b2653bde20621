// linear.cuh -- the projections either side of the joint as tcgen05 GEMMs.
//
// Reference: speechbrain.nnet.linear.Linear.forward (SB/nnet/linear.py:63-76, nn.Linear at :61) as instantiated for
// encoder_proj [B*T,256] x [256,640] and decoder_proj [B*U,512] x [512,640] (hparams/LibriSpeechMix/conformer-t_scratch.yaml:
// 172-174,187-189; called at train_librispeechmix_scratch.py:122,127), and autograd's backward of it (dX = dY W,
// dW = dY^T X, db = sum_r dY).  The reference runs them as fp32 cuBLAS SIMT GEMMs and hands fp32 [B,T,640] / [B,U,640]
// tensors to the joiner; here ONE kernel template serves all three products
//
//      C[m, n] = sum_k A(m, k) * B(n, k)            A(m,k) = A[m*a_ld_mn + k*a_ld_k],  B likewise (one stride is 1)
//
// with fp32 operands read straight from global memory -- the joint backward's d_enc / d_dec are consumed as they are --
// and split IN THE KERNEL into bf16 (hi, lo) pairs written to K-major SWIZZLE_128B shared-memory images.  Three MMAs per
// K = 16 step (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) keep 16-17 bits per term (what is dropped -- lo*lo and the
// rounding of each lo part -- is at most 3 * 2^-16 of |x||w| per term and averages out over the contraction: measured
// 1.4e-5 of the largest output at K = 256, against 4e-3 for plain bf16 operands), far below the bf16 rounding the joint
// applies to these outputs anyway, although the products run on the bf16 tensor cores; at
// 2-3 GFLOP per product the tripled MMA count is noise.  The forward epilogue adds the bias and writes the fp32 result
// AND its bf16 rounding in one pass: the bf16 copy is the operand image the joint GEMM's producers read, so the separate
// fp32 -> bf16 pass over enc_out / dec_out disappears.
//
// Tile: 128 (m) x 128 (n) x 64 (k) per k-block, 256 threads, 128 registers, one 68 KB operand stage and 256 TMEM columns
// per CTA so that TWO CTAs share an SM, cta_group::1 MMAs (M = 128, N = 128, or N = 144 for the tile that carries the
// "ones" row, see below).  All 8 warps load and convert (the global loads of k-block i+1 are issued right after block i
// has been written to shared memory, so they fly while its MMAs run); warp 0 issues the MMAs; all 8 warps run the epilogue
// (TMEM -> registers -> a padded shared-memory staging tile -> coalesced 256-byte row segments).  A CTA's k-blocks are a
// latency chain (load -> convert -> MMA); the second resident CTA fills its gaps (measured: one CTA per SM with two stages
// 30.8 us for encoder_proj's forward, this layout see profiles/r2_linear.txt).
//
// dW is a [N_out, K_in] product contracted over the R rows: split-K over gridDim.z with fp32 partials folded in a fixed
// order by linear_fold_kernel (deterministic, no atomics).  db rides along as ONE extra B row of ones appended to the last
// n tile (UMMA N = 144): column N of the accumulator is sum_k A(m, k) = the column sum of dY.
#pragma once

#include "common.cuh"

namespace tsasr {

static constexpr int kLinThreads = 256;
static constexpr int kLinBM = 128, kLinBN = 128, kLinBK = 64, kLinExtraN = 16;
static constexpr uint32_t kLinAImg = kLinBM * 128;                       // 128 rows x 64 bf16 = 16 KB
static constexpr uint32_t kLinBImg = (kLinBN + kLinExtraN) * 128;        // 144 rows = 18 KB
static constexpr uint32_t kLinStage = 2 * kLinAImg + 2 * kLinBImg;       // A hi, A lo, B hi, B lo = 68 KB
static constexpr int kLinStages = 1;   // one operand stage per CTA, TWO CTAs per SM: the k-blocks of a CTA are latency-bound (global
                                       // load -> convert -> MMA), a second resident CTA fills the gaps better than a second stage
static constexpr uint32_t kLinBarOff = kLinStages * kLinStage;
static constexpr uint32_t kLinSmemBytes = kLinBarOff + 64;
static constexpr int kLinStgLd = 68;                                     // floats per staged row (64 + 4: conflict-free float4)
static_assert(8 * 32 * kLinStgLd * 4 <= kLinStages * kLinStage, "epilogue staging must fit the operand stages");

struct LinParams {
    const float* A; long long a_ld_mn, a_ld_k;
    const float* B; long long b_ld_mn, b_ld_k;
    const float* bias;              // [N] added along n, or nullptr
    float* C32;                     // [M, c_ld] (+ blockIdx.z * c_split_stride), or nullptr
    __nv_bfloat16* C16;             // same layout, or nullptr
    long long c_ld, c_split_stride;
    float* ones_out;                // [gridDim.z][M]: sum_k A(m, k) (the "ones" row), or nullptr
    int M, N, K;
    int kb_per_split;               // k-blocks (of 64) per gridDim.z slice; every slice is non-empty
    int a_vec, b_vec, c_vec;        // 16-byte vector paths allowed (alignment checked on the host)
};

// one operand tile (128 rows x 64 k) -> 32 registers per thread
//   KCONTIG : chunk q = tid + 256 i -> row q >> 3, 16-byte chunk q & 7 (8 consecutive k): lanes 0-7 read one 256-byte row segment
//   !KCONTIG: row = tid & 127 (consecutive lanes = consecutive rows = consecutive addresses), chunk (tid >> 7) + 2 i
template <bool KCONTIG>
__device__ __forceinline__ void lin_load_tile(const float* __restrict__ base, long long ld_mn, long long ld_k, int mn0, int MN, int k0,
                                              int K, int vec, float (&r)[32]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int row, c;
        if (KCONTIG) { const int q = tid + 256 * i; row = q >> 3; c = q & 7; }
        else { row = tid & 127; c = (tid >> 7) + 2 * i; }
        const int mn = mn0 + row, k = k0 + c * 8;
        if (KCONTIG) {
            const float* src = base + (long long)mn * ld_mn + k;
            if (mn < MN && k + 8 <= K && vec) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                r[8 * i + 0] = v0.x; r[8 * i + 1] = v0.y; r[8 * i + 2] = v0.z; r[8 * i + 3] = v0.w;
                r[8 * i + 4] = v1.x; r[8 * i + 5] = v1.y; r[8 * i + 6] = v1.z; r[8 * i + 7] = v1.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) r[8 * i + j] = (mn < MN && k + j < K) ? __ldg(src + j) : 0.f;
            }
        } else {
            const float* src = base + (long long)k * ld_k + (long long)mn * ld_mn;
#pragma unroll
            for (int j = 0; j < 8; ++j) r[8 * i + j] = (mn < MN && k + j < K) ? __ldg(src + (long long)j * ld_k) : 0.f;
        }
    }
}

// registers -> (hi, lo) bf16 images, K-major SWIZZLE_128B
template <bool KCONTIG>
__device__ __forceinline__ void lin_store_tile(uint8_t* img_hi, uint8_t* img_lo, const float (&r)[32]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int row, c;
        if (KCONTIG) { const int q = tid + 256 * i; row = q >> 3; c = q & 7; }
        else { row = tid & 127; c = (tid >> 7) + 2 * i; }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            // packed conversions only (F2FP.BF16.PACK_AB, ALU pipe): the scalar cvt is an XU-pipe F2F at a quarter of the rate
            const float x0 = r[8 * i + 2 * w], x1 = r[8 * i + 2 * w + 1];
            hi[w] = pack_bf16x2(x0, x1);
            lo[w] = pack_bf16x2(x0 - bf16_lo(hi[w]), x1 - bf16_hi(hi[w]));
        }
        const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)c);
        *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(kLinThreads, 2) linear_gemm_kernel(const LinParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* stage_free = reinterpret_cast<uint64_t*>(smem + kLinBarOff);  // [2]
    uint64_t* acc_full = stage_free + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(stage_free + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kLinBM, n0 = blockIdx.y * kLinBN;
    const bool ones_tile = p.ones_out != nullptr && blockIdx.y == gridDim.y - 1;
    const int kb_total = (p.K + kLinBK - 1) / kLinBK;
    const int kb_begin = blockIdx.z * p.kb_per_split, kb_end = min(kb_total, kb_begin + p.kb_per_split);

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B atoms need 1024-byte alignment
        mbar_init(&stage_free[0], 1);
        mbar_init(&stage_free[1], 1);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_ptr);
    if (ones_tile) {
        // rows 128..143 of the B images never change: row 128 of "hi" is all ones (every 16-byte chunk holds the same
        // pattern, so the swizzle is irrelevant), the rest is zero
        for (int i = threadIdx.x; i < kLinStages * 2 * 128; i += kLinThreads) {
            const int s = i >> 8, img = (i >> 7) & 1, chunk = i & 127;  // 128 chunks of 16 bytes = rows 128..143
            const uint32_t v = (img == 0 && chunk < 8) ? 0x3F803F80u : 0u;
            *reinterpret_cast<uint4*>(smem + s * kLinStage + 2 * kLinAImg + img * kLinBImg + kLinBN * 128 + chunk * 16) = make_uint4(v, v, v, v);
        }
        fence_proxy_async_smem();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t idesc = make_idesc_bf16(kLinBM, ones_tile ? kLinBN + kLinExtraN : kLinBN, 0, 0);

    float ra[32], rb[32];
    lin_load_tile<A_KCONTIG>(p.A, p.a_ld_mn, p.a_ld_k, m0, p.M, kb_begin * kLinBK, p.K, p.a_vec, ra);
    lin_load_tile<B_KCONTIG>(p.B, p.b_ld_mn, p.b_ld_k, n0, p.N, kb_begin * kLinBK, p.K, p.b_vec, rb);
    uint32_t it = 0;
#pragma unroll 1
    for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const uint32_t s = it % kLinStages, use = it / kLinStages;
        uint8_t* st = smem + s * kLinStage;
        // the MMAs that read this stage kLinStages k-blocks ago have completed
        if (use >= 1) mbar_wait(&stage_free[s], (use - 1) & 1, 0x900 | s);
        lin_store_tile<A_KCONTIG>(st, st + kLinAImg, ra);
        lin_store_tile<B_KCONTIG>(st + 2 * kLinAImg, st + 2 * kLinAImg + kLinBImg, rb);
        if (kb + 1 < kb_end) {  // next k-block: in flight while this one is multiplied
            lin_load_tile<A_KCONTIG>(p.A, p.a_ld_mn, p.a_ld_k, m0, p.M, (kb + 1) * kLinBK, p.K, p.a_vec, ra);
            lin_load_tile<B_KCONTIG>(p.B, p.b_ld_mn, p.b_ld_k, n0, p.N, (kb + 1) * kLinBK, p.K, p.b_vec, rb);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {  // whole warp, converged; one lane is elected inside each issuing instruction
            tcgen05_fence_after();
            const uint32_t sb = smem_u32(st);
            const uint64_t a_hi = make_smem_desc_sw128(sb, 0, 1024), a_lo = make_smem_desc_sw128(sb + kLinAImg, 0, 1024);
            const uint64_t b_hi = make_smem_desc_sw128(sb + 2 * kLinAImg, 0, 1024);
            const uint64_t b_lo = make_smem_desc_sw128(sb + 2 * kLinAImg + kLinBImg, 0, 1024);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // 32 bytes (2 x 16-byte units) per K = 16 step inside the swizzled row
                umma_bf16_e(tmem_base, a_hi + 2 * ks, b_hi + 2 * ks, idesc, (it | (uint32_t)ks) != 0u);
                umma_bf16_e(tmem_base, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
                umma_bf16_e(tmem_base, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
            }
            umma_commit_e(&stage_free[s]);
            if (kb + 1 == kb_end) umma_commit_e(acc_full);
        }
    }
    mbar_wait(acc_full, 0, 0x910);
    tcgen05_fence_after();

    // ---- epilogue: warp w owns TMEM lanes 32 (w & 3) .. +31 and columns 64 (w >> 2) .. +63 ----
    const int q = warp & 3, half = warp >> 2;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stg = reinterpret_cast<float*>(smem) + warp * (32 * kLinStgLd);
    {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(trow + half * 64, r0);
        tmem_ld_32x32b_x32(trow + half * 64 + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            *reinterpret_cast<uint4*>(stg + lane * kLinStgLd + j) = make_uint4(r0[j], r0[j + 1], r0[j + 2], r0[j + 3]);
            *reinterpret_cast<uint4*>(stg + lane * kLinStgLd + 32 + j) = make_uint4(r1[j], r1[j + 1], r1[j + 2], r1[j + 3]);
        }
    }
    __syncwarp();
    float* c32 = p.C32 ? p.C32 + (long long)blockIdx.z * p.c_split_stride : nullptr;
    __nv_bfloat16* c16 = p.C16;
    const int c4 = (lane & 15) * 4, n = n0 + half * 64 + c4;  // a lane keeps its four columns for all 16 row pairs
    const bool vec = p.c_vec && n + 4 <= p.N;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias && vec) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll 4
    for (int rr = 0; rr < 32; rr += 2) {
        const int rl = rr + (lane >> 4);
        const int m = m0 + q * 32 + rl;
        if (m >= p.M || n >= p.N) continue;
        float4 v = *reinterpret_cast<const float4*>(stg + rl * kLinStgLd + c4);
        const long long o = (long long)m * p.c_ld + n;
        if (vec) {
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            if (c32) *reinterpret_cast<float4*>(c32 + o) = v;
            if (c16) *reinterpret_cast<uint2*>(c16 + o) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        } else {
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n + j < p.N) {
                    const float y = e[j] + (p.bias ? __ldg(p.bias + n + j) : 0.f);
                    if (c32) c32[o + j] = y;
                    if (c16) c16[o + j] = __float2bfloat16_rn(y);
                }
            }
        }
    }
    if (ones_tile && half == 1) {  // accumulator column 128 = the product with the row of ones
        uint32_t rx[16];
        tmem_ld_32x32b_x16(trow + kLinBN, rx);
        tmem_ld_wait();
        const int m = m0 + q * 32 + lane;
        if (m < p.M) p.ones_out[(long long)blockIdx.z * p.M + m] = __uint_as_float(rx[0]);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// out[i] = sum_z partial[z * split_stride + i] (fixed order), and -- when ones_partial is given -- db[m] = sum_z ones_partial[z * M + m]
__global__ void __launch_bounds__(256)
linear_fold_kernel(const float* __restrict__ partial, long long split_stride, int splits, long long n, float* __restrict__ out,
                   const float* __restrict__ ones_partial, int M, float* __restrict__ db) {
    const long long stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((n & 3) == 0) {
        for (long long i = tid; i < (n >> 2); i += stride) {
            float4 acc = __ldg(reinterpret_cast<const float4*>(partial) + i);
            for (int z = 1; z < splits; ++z) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (long long)z * split_stride) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            reinterpret_cast<float4*>(out)[i] = acc;
        }
    } else {
        for (long long i = tid; i < n; i += stride) {
            float acc = partial[i];
            for (int z = 1; z < splits; ++z) acc += partial[(long long)z * split_stride + i];
            out[i] = acc;
        }
    }
    if (ones_partial && db) {
        for (long long i = tid; i < M; i += stride) {
            float acc = ones_partial[i];
            for (int z = 1; z < splits; ++z) acc += ones_partial[(long long)z * M + i];
            db[i] = acc;
        }
    }
}

}  // namespace tsasr
