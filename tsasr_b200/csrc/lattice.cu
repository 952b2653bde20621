// lattice.cu -- RNN-T lattice kernels (HBM/latency-bound part of the hot path).
//
//   logits_to_lattice_kernel : [B,T,U,V] logits -> 2-value lattice {lp_blank, lp_emit} + log-sum-exp
//                              (replaces torchaudio's ReduceMax2D / ReduceLogSumExpGivenMax2D /
//                               ComputeLogProbs and SB/nnet/losses.py:84 log_softmax on the compat path)
//   alpha_beta_kernel        : anti-diagonal wavefront forward/backward DP
//                              (replaces SB/nnet/loss/transducer_loss.py:31-180 cu_kernel_forward /
//                               cu_kernel_backward and torchaudio's ComputeAlphasBetasCosts)
//   logits_grad_kernel       : dense d cost / d logits with the softmax folded in (torchaudio
//                              ComputeGradients semantics)
//   logprobs_grad_kernel     : sparse gradient w.r.t. log-probs (SB/nnet/loss/transducer_loss.py:183-236)
//
// All lattice-sized arrays use the skewed layout of common.cuh (diagonal-major, u contiguous).
#include "common.cuh"

#include <cuda_fp16.h>

namespace tsasr {

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr float kLn2 = 0.6931471805599453f;

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u4(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
// 8 half-precision values of one 16-byte word <-> 8 floats
template <typename T>
__device__ __forceinline__ void unpack8(const uint4& w, float* x) {
    const T* h = reinterpret_cast<const T*>(&w);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = to_float<T>(h[i]);
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float* x) {
    uint4 w;
    T* h = reinterpret_cast<T*>(&w);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = from_float<T>(x[i]);
    return w;
}

// online (max, sum) update with a block of values already in registers
__device__ __forceinline__ void online_update(float& m, float& s, const float* x, int n) {
    float cm = x[0];
#pragma unroll
    for (int i = 1; i < 32; ++i)
        if (i < n) cm = fmaxf(cm, x[i]);
    const float mn = fmaxf(m, cm);
    if (mn == -INFINITY) return;  // nothing but -inf so far
    const float ms = mn * kLog2e;
    float acc = s * exp2f(m * kLog2e - ms);
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) acc += exp2f(fmaf(x[i], kLog2e, -ms));
    s = acc;
    m = mn;
}

__device__ __forceinline__ void warp_combine(float& m, float& s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const float mn = fmaxf(m, m2);
        if (mn == -INFINITY) { s = 0.f; m = mn; continue; }
        s = s * exp2f((m - mn) * kLog2e) + s2 * exp2f((m2 - mn) * kLog2e);
        m = mn;
    }
}

// One warp per lattice cell (row of V logits).  Rows outside the utterance's T_b x U_b rectangle are
// skipped entirely (no HBM traffic).  normalized != 0: input rows are already log-probs (den := 0).
template <typename T>
__global__ void __launch_bounds__(256)
logits_to_lattice_kernel(const T* __restrict__ logits, const int* __restrict__ targets,
                         const int* __restrict__ logit_lengths, const int* __restrict__ target_lengths,
                         int B, int Tmax, int U, int V, int blank, int normalized,
                         float2* __restrict__ lat2, float* __restrict__ den) {
    const int warps_per_block = blockDim.x >> 5;
    const long long cell = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (cell >= (long long)B * Tmax * U) return;
    const int u = (int)(cell % U);
    const int t = (int)((cell / U) % Tmax);
    const int b = (int)(cell / ((long long)U * Tmax));
    int Tb, Ub;
    clamped_lengths(logit_lengths, target_lengths, b, Tmax, U, Tb, Ub);  // same rectangle as the DP (common.cuh)
    if (t >= Tb || u >= Ub) return;
    const T* row = logits + (size_t)cell * V;

    float lse = 0.f;
    if (!normalized) {
        float m = -INFINITY, s = 0.f;
        if (sizeof(T) == 4 && (V & 3) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            const int V4 = V >> 2;
            for (int base = 0; base < V4; base += 32 * 8) {
                float x[32];
                int n = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int idx = base + j * 32 + lane;
                    if (idx < V4) {
                        const float4 v = ldg_stream_f4(row4 + idx);
                        x[4 * j + 0] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
                        n = 4 * j + 4;
                    } else {
                        x[4 * j + 0] = x[4 * j + 1] = x[4 * j + 2] = x[4 * j + 3] = -INFINITY;
                    }
                }
                if (n > 0) online_update(m, s, x, 32);
            }
        } else if (sizeof(T) == 2 && (V & 7) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
            // half-precision rows: 16-byte streaming loads, 8 values each
            const uint4* row8 = reinterpret_cast<const uint4*>(row);
            const int V8 = V >> 3;
            for (int base = 0; base < V8; base += 32 * 4) {
                float x[32];
                int n = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = base + j * 32 + lane;
                    if (idx < V8) {
                        unpack8<T>(ldg_stream_u4(row8 + idx), x + 8 * j);
                        n = 8 * j + 8;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[8 * j + i] = -INFINITY;
                    }
                }
                if (n > 0) online_update(m, s, x, 32);
            }
        } else {
            for (int base = 0; base < V; base += 32 * 32) {
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int idx = base + j * 32 + lane;
                    x[j] = idx < V ? to_float<T>(row[idx]) : -INFINITY;
                }
                online_update(m, s, x, 32);
            }
        }
        warp_combine(m, s);
        lse = m + __logf(s);
    }
    if (lane == 0) {
        const float xb = to_float<T>(row[blank]);
        float xe = -INFINITY;
        if (u < Ub - 1) xe = to_float<T>(row[min(max(targets[(size_t)b * (U - 1) + u], 0), V - 1)]) - lse;  // label ids clamped: never read outside the row
        const size_t o = skew_index(b, t, u, Tmax, U);
        lat2[o] = make_float2(xb - lse, xe);
        den[o] = lse;
    }
}

// ------------------------------------------------------------------------------------------------
// Wavefront DP.  grid = (B, 2): blockIdx.y = 0 -> alpha, 1 -> beta.  Thread j owns one lattice
// column (alpha: u = j, beta: u = U_b-1-j) and walks it in time; step s touches the anti-diagonal
// d = s (alpha) or d = T_b+U_b-2-s (beta), which is one contiguous row of the skewed layout.  The
// neighbour value moves by warp shuffle; between warps through a double-buffered shared slot and
// one named barrier per step.  Lattice rows are prefetched PF diagonals ahead (they do not depend
// on the DP state), so the per-step critical path is shuffle + logaddexp only.
// ------------------------------------------------------------------------------------------------
static constexpr int kDpPrefetch = 8;       // diagonals in flight per column (cp.async ring in shared memory); 8 measured best of 4..16
static constexpr float kDpNeg = -1.0e30f;  // "log 0" sentinel: finite, so no -inf special cases on the chain

// log2(2^a + 2^b): the whole per-step dependency chain is FADD -> MUFU.EX2 -> FADD -> MUFU.LG2 -> FADD
__device__ __forceinline__ float logaddexp2_fast(float a, float b) {
    return fmaxf(a, b) + lg2_approx(1.f + ex2_approx(-fabsf(a - b)));
}

// One in-order warp per 32 columns walks ~T+U dependent steps, so the per-step critical path is all that
// matters.  Lattice values are staged by per-thread cp.async into a shared-memory ring kDpPrefetch
// diagonals ahead (cp.async groups retire in order, so `wait_group kDpPrefetch-1` waits for exactly the
// oldest one; register prefetch cannot do that: scoreboard slots are shared between the loads in flight),
// loop invariants are pinned in registers, and the boundary cases are folded into the initial values.
template <bool BETA>
__device__ __forceinline__ void dp_pass(const float2* __restrict__ lat, float* __restrict__ out, int Tb, int Ub,
                                        int U, float* __restrict__ result, float* xchg /* [2][32] */, float2* ring) {
    int j;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(j));
    const int lane = j & 31, warp = j >> 5;
    const int nwarps = (Ub + 31) >> 5;
    if (warp >= nwarps) return;
    const int nthreads_active = nwarps << 5;
    const bool col_ok = j < Ub;
    const unsigned Tb_eff = col_ok ? (unsigned)Tb : 0u;  // (unsigned)tl < Tb_eff  <=>  column active at tl
    const int u = col_ok ? (BETA ? (Ub - 1 - j) : j) : 0;
    const int S = Tb + Ub - 1;
    const long long stride = BETA ? -(long long)U : (long long)U;
    const float2* lp = lat + (BETA ? (long long)(S - 1) * U : 0ll) + u;
    float* op = out + (BETA ? (long long)(S - 1) * U : 0ll) + u;
    // shared exchange slots: warp w writes slot [parity][w], warp w+1 reads it one step later
    uint32_t x_wr = smem_u32(xchg) + warp * 4, x_rd = x_wr - 4;
    uint32_t r_base = smem_u32(ring) + j * 8;           // this column's ring: slot k at r_base + k * slot_stride
    const uint32_t slot_stride = blockDim.x * 8;
    asm volatile("" : "+r"(x_rd), "+r"(x_wr), "+r"(r_base));
    const bool rd_lds = lane == 0 && warp > 0;
    const bool wr_sts = lane == 31;
    const bool multi = nwarps > 1;

    int tl = -j;  // local time of this column at step s
#pragma unroll
    for (int k = 0; k < kDpPrefetch; ++k) {
        if ((unsigned)(tl + k) < Tb_eff)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(r_base + k * slot_stride), "l"(lp) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        lp += stride;
    }

    // alpha: carry = a(t-1,u)+blank(t-1,u), pass = a(t,u)+emit(t,u); beta: carry = pass = b(t,u) (log2 units).
    // Column 0 starts from carry = 0 so the generic update yields a(0,0) = 0 / b(T-1,U-1) = lp_blank.
    float carry = j == 0 ? 0.f : kDpNeg;
    float pass = kDpNeg;
    float res = 0.f;

    for (int s0 = 0; s0 < S; s0 += kDpPrefetch) {
#pragma unroll
        for (int k = 0; k < kDpPrefetch; ++k) {
            if (s0 + k < S) {  // block-uniform
                asm volatile("cp.async.wait_group %0;" ::"n"(kDpPrefetch - 1) : "memory");
                float2 raw;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(raw.x), "=f"(raw.y) : "r"(r_base + k * slot_stride));
                if ((unsigned)(tl + kDpPrefetch) < Tb_eff)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(r_base + k * slot_stride), "l"(lp) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                lp += stride;
                const float2 cell = make_float2(raw.x * kLog2e, raw.y * kLog2e);
                float from_left = __shfl_up_sync(0xffffffffu, pass, 1);
                if (lane == 0) from_left = kDpNeg;
                if (rd_lds) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(from_left) : "r"(x_rd + ((k + 1) & 1) * 128));
                const bool active = (unsigned)tl < Tb_eff;
                float val;
                if (BETA) {
                    val = logaddexp2_fast(carry + cell.x, from_left + cell.y);
                    if (active) { carry = val; pass = val; res = val; }
                } else {
                    val = logaddexp2_fast(carry, from_left);
                    if (active) { carry = val + cell.x; pass = val + cell.y; res = carry; }
                }
                if (active) *op = val * kLn2;
                op += stride;
                ++tl;
                if (multi) {
                    if (wr_sts) asm volatile("st.shared.f32 [%0], %1;" ::"r"(x_wr + (k & 1) * 128), "f"(pass) : "memory");
                    asm volatile("bar.sync 1, %0;" ::"r"(nthreads_active) : "memory");
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (j == Ub - 1) *result = res * kLn2;  // log P(y|x): last active step of the last column
}

__global__ void __launch_bounds__(1024)
alpha_beta_kernel(const float2* __restrict__ lat2, const int* __restrict__ logit_lengths,
                  const int* __restrict__ target_lengths, int Tmax, int U, float* __restrict__ alpha,
                  float* __restrict__ beta, float* __restrict__ ll_alpha, float* __restrict__ ll_beta) {
    __shared__ float xchg[64];
    extern __shared__ __align__(16) float2 dp_ring[];  // [kDpPrefetch][blockDim.x]
    griddep_wait();  // lat2 comes from the preceding kernel (programmatic dependent launch)
    const int b = blockIdx.x;
    // lengths are clamped to the padded lattice: the fused path validates them on the host only after
    // launching (tsasr_b200/functional.py), so an invalid length must never index outside the slab
    int Tb, Ub;
    clamped_lengths(logit_lengths, target_lengths, b, Tmax, U, Tb, Ub);
    const size_t base = (size_t)b * (size_t)(Tmax + U - 1) * (size_t)U;
    // Each utterance's own rectangle starts at diagonal 0 of its slab: cell (t,u) -> row t+u.
    if (blockIdx.y == 0) dp_pass<false>(lat2 + base, alpha + base, Tb, Ub, U, ll_alpha + b, xchg, dp_ring);
    else dp_pass<true>(lat2 + base, beta + base, Tb, Ub, U, ll_beta + b, xchg, dp_ring);
}

// Wide lattices (U > 1024 columns: more than one thread block of columns).  Same recursion, simplest possible schedule:
// one CTA per (utterance, direction), the previous anti-diagonal kept in shared memory (double buffered), every thread
// takes the columns j, j + blockDim, ... of the current diagonal, one __syncthreads per diagonal.  Not tuned -- an
// utterance with more than 1023 labels is far outside the recipe's range -- but it lifts the limit the reference's
// torchaudio path does not have (its Numba path does: one thread per column as well).
template <bool BETA>
__device__ __forceinline__ void dp_pass_wide(const float2* __restrict__ lat, float* __restrict__ out, int Tb, int Ub, int U,
                                             float* __restrict__ result, float* diag /* [2][U] */) {
    const int S = Tb + Ub - 1;  // diagonals of the utterance's own rectangle
    float* prev = diag;
    float* cur = diag + U;
    for (int s = 0; s < S; ++s) {
        // alpha walks d = s upwards from cell (0,0); beta walks the mirrored lattice (t' = Tb-1-t, u' = Ub-1-u) the same way
        for (int uu = threadIdx.x; uu < Ub; uu += blockDim.x) {
            const int tt = s - uu;
            if (tt < 0 || tt >= Tb) continue;
            const int t = BETA ? Tb - 1 - tt : tt, u = BETA ? Ub - 1 - uu : uu;
            const size_t o = (size_t)(t + u) * U + u;
            float val;
            if (BETA) {
                const float2 cell = lat[o];
                if (s == 0) {
                    val = cell.x;  // beta(T-1, U-1) = lp_blank
                } else {
                    const float a = tt > 0 ? prev[uu] + cell.x : -INFINITY;        // beta(t+1, u) + lp_blank(t, u)
                    const float b = uu > 0 ? prev[uu - 1] + cell.y : -INFINITY;    // beta(t, u+1) + lp_emit(t, u)
                    val = logaddexp_fast(a, b);
                }
            } else {
                if (s == 0) {
                    val = 0.f;     // alpha(0, 0) = 0
                } else {
                    const float a = tt > 0 ? prev[uu] + lat[(size_t)(t - 1 + u) * U + u].x : -INFINITY;          // via blank from (t-1, u)
                    const float b = uu > 0 ? prev[uu - 1] + lat[(size_t)(t + u - 1) * U + (u - 1)].y : -INFINITY;  // via emit from (t, u-1)
                    val = logaddexp_fast(a, b);
                }
            }
            cur[uu] = val;
            out[o] = val;
            if (s == S - 1) *result = BETA ? val : val + lat[o].x;  // log P: beta(0,0), or alpha(T-1,U-1) + lp_blank(T-1,U-1)
        }
        __syncthreads();
        float* tmp = prev; prev = cur; cur = tmp;
    }
}

__global__ void __launch_bounds__(1024)
alpha_beta_wide_kernel(const float2* __restrict__ lat2, const int* __restrict__ logit_lengths,
                       const int* __restrict__ target_lengths, int Tmax, int U, float* __restrict__ alpha,
                       float* __restrict__ beta, float* __restrict__ ll_alpha, float* __restrict__ ll_beta) {
    extern __shared__ __align__(16) float dp_diag[];  // [2][U]
    griddep_wait();
    const int b = blockIdx.x;
    int Tb, Ub;
    clamped_lengths(logit_lengths, target_lengths, b, Tmax, U, Tb, Ub);
    const size_t base = (size_t)b * (size_t)(Tmax + U - 1) * (size_t)U;
    if (blockIdx.y == 0) dp_pass_wide<false>(lat2 + base, alpha + base, Tb, Ub, U, ll_alpha + b, dp_diag);
    else dp_pass_wide<true>(lat2 + base, beta + base, Tb, Ub, U, ll_beta + b, dp_diag);
}

// cost[b] = -log P from the beta pass (torchaudio: costs = -beta(0,0)); written by a tiny kernel so the
// DP kernel keeps both log-likelihoods available for the consistency check in tests.
__global__ void finalize_cost_kernel(const float* __restrict__ ll_beta, float* __restrict__ cost, int B) {
    griddep_wait();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) cost[b] = -ll_beta[b];
}

// Per-cell occupation terms shared by every gradient kernel:
//   occ = dy * exp(a + b - L)                      (= ob + oe)
//   ob  = dy * exp(a + lp_blank + b(t+1,u) - L)    (terminal blank: b := 0)
//   oe  = dy * exp(a + lp_emit  + b(t,u+1) - L)    (u < U_b-1)
struct CellTerms { float occ, ob, oe; };
__device__ __forceinline__ CellTerms cell_terms(const float2* __restrict__ lat2, const float* __restrict__ alpha,
                                                const float* __restrict__ beta, float L, float dy, int b, int t, int u,
                                                int Tb, int Ub, int Tmax, int U) {
    const size_t o = skew_index(b, t, u, Tmax, U);
    const float a = alpha[o];
    const float2 lp = lat2[o];
    CellTerms r;
    r.occ = dy * __expf(a + beta[o] - L);
    float bt1 = -INFINITY;
    if (t < Tb - 1) bt1 = beta[o + U];  // (t+1,u): next diagonal, same u
    else if (u == Ub - 1) bt1 = 0.f;
    r.ob = bt1 == -INFINITY ? 0.f : dy * __expf(a + lp.x + bt1 - L);
    r.oe = u < Ub - 1 ? dy * __expf(a + lp.y + beta[o + U + 1] - L) : 0.f;  // (t,u+1): next diagonal, u+1
    return r;
}

// Dense d cost/d logits, one warp per cell; cells outside the rectangle get exact zeros.
template <typename T>
__global__ void __launch_bounds__(256)
logits_grad_kernel(const T* __restrict__ logits, const int* __restrict__ targets,
                   const int* __restrict__ logit_lengths, const int* __restrict__ target_lengths, int B, int Tmax,
                   int U, int V, int blank, const float2* __restrict__ lat2, const float* __restrict__ den,
                   const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ cost,
                   const float* __restrict__ dcost, float clamp, T* __restrict__ dlogits) {
    const int warps_per_block = blockDim.x >> 5;
    const long long cell = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (cell >= (long long)B * Tmax * U) return;
    const int u = (int)(cell % U);
    const int t = (int)((cell / U) % Tmax);
    const int b = (int)(cell / ((long long)U * Tmax));
    int Tb, Ub;
    clamped_lengths(logit_lengths, target_lengths, b, Tmax, U, Tb, Ub);  // same rectangle as the DP (common.cuh)
    const T* row = logits + (size_t)cell * V;
    T* out = dlogits + (size_t)cell * V;
    const bool vec = sizeof(T) == 4 && (V & 3) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const bool vec8 = sizeof(T) == 2 && (V & 7) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (t >= Tb || u >= Ub) {
        if (vec) {
            float4* o4 = reinterpret_cast<float4*>(out);
            for (int i = lane; i < (V >> 2); i += 32) stg_stream_f4(o4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
        } else if (vec8) {
            uint4* o8 = reinterpret_cast<uint4*>(out);
            for (int i = lane; i < (V >> 3); i += 32) stg_stream_u4(o8 + i, make_uint4(0u, 0u, 0u, 0u));
        } else {
            for (int i = lane; i < V; i += 32) out[i] = from_float<T>(0.f);
        }
        return;
    }
    // torchaudio's order (ComputeGradients, then `grad * dy` in the autograd backward, functional.py:1729-1734): the
    // gradient of the UNIT cost is clamped, the upstream factor dcost[b] (1/B under reduction="mean") multiplies after.
    const float dy = dcost ? dcost[b] : 1.f;
    const CellTerms ct = cell_terms(lat2, alpha, beta, -cost[b], 1.f, b, t, u, Tb, Ub, Tmax, U);
    const float nd = -den[skew_index(b, t, u, Tmax, U)] * kLog2e;
    const int label = u < Ub - 1 ? targets[(size_t)b * (U - 1) + u] : -1;
    auto grad = [&](float x, int v) {
        float g = exp2f(fmaf(x, kLog2e, nd)) * ct.occ;
        if (v == blank) g -= ct.ob;
        if (v == label) g -= ct.oe;
        if (clamp > 0.f) g = fminf(fmaxf(g, -clamp), clamp);
        return g * dy;
    };
    if (vec) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        float4* o4 = reinterpret_cast<float4*>(out);
        const int V4 = V >> 2;
        for (int base = 0; base < V4; base += 32 * 4) {
            float4 x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = base + j * 32 + lane;
                if (idx < V4) x[j] = ldg_stream_f4(r4 + idx);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = base + j * 32 + lane;
                if (idx < V4) {
                    float4 g;
                    g.x = grad(x[j].x, 4 * idx + 0);
                    g.y = grad(x[j].y, 4 * idx + 1);
                    g.z = grad(x[j].z, 4 * idx + 2);
                    g.w = grad(x[j].w, 4 * idx + 3);
                    stg_stream_f4(o4 + idx, g);
                }
            }
        }
    } else if (vec8) {
        const uint4* r8 = reinterpret_cast<const uint4*>(row);
        uint4* o8 = reinterpret_cast<uint4*>(out);
        const int V8 = V >> 3;
        for (int base = 0; base < V8; base += 32 * 4) {
            uint4 w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = base + j * 32 + lane;
                if (idx < V8) w[j] = ldg_stream_u4(r8 + idx);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = base + j * 32 + lane;
                if (idx < V8) {
                    float x[8];
                    unpack8<T>(w[j], x);
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = grad(x[i], 8 * idx + i);
                    stg_stream_u4(o8 + idx, pack8<T>(x));
                }
            }
        }
    } else {
        for (int i = lane; i < V; i += 32) out[i] = from_float<T>(grad(to_float<T>(row[i]), i));
    }
}

// Sparse gradient w.r.t. log-probs (Numba semantics): -ob at blank, -oe at labels[b,u]; the dense
// zero fill is a cudaMemsetAsync issued by the caller.  One thread per cell.
__global__ void __launch_bounds__(256)
logprobs_grad_kernel(const int* __restrict__ targets, const int* __restrict__ logit_lengths,
                     const int* __restrict__ target_lengths, int B, int Tmax, int U, int V, int blank,
                     const float2* __restrict__ lat2, const float* __restrict__ alpha,
                     const float* __restrict__ beta, const float* __restrict__ cost,
                     const float* __restrict__ dcost, float* __restrict__ grads) {
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (long long)B * Tmax * U) return;
    const int u = (int)(cell % U);
    const int t = (int)((cell / U) % Tmax);
    const int b = (int)(cell / ((long long)U * Tmax));
    int Tb, Ub;
    clamped_lengths(logit_lengths, target_lengths, b, Tmax, U, Tb, Ub);  // same rectangle as the DP (common.cuh)
    if (t >= Tb || u >= Ub) return;
    const float dy = dcost ? dcost[b] : 1.f;
    const CellTerms ct = cell_terms(lat2, alpha, beta, -cost[b], dy, b, t, u, Tb, Ub, Tmax, U);
    float* g = grads + (size_t)cell * V;
    if (t < Tb - 1 || u == Ub - 1) g[blank] = -ct.ob;
    if (u < Ub - 1) g[min(max(targets[(size_t)b * (U - 1) + u], 0), V - 1)] = -ct.oe;  // label ids clamped: never write outside the row
}

// ------------------------------------------------------------------------------------------------
// host-side launchers (called from capi.cu)
// ------------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t launch_logits_to_lattice_t(const void* logits, const int* targets, const int* ll, const int* tl,
                                              int B, int Tmax, int U, int V, int blank, int normalized, float2* lat2,
                                              float* den, cudaStream_t st) {
    const long long cells = (long long)B * Tmax * U;
    const int wpb = 8;
    const long long blocks = (cells + wpb - 1) / wpb;
    logits_to_lattice_kernel<T><<<(unsigned)blocks, wpb * 32, 0, st>>>(static_cast<const T*>(logits), targets, ll, tl,
                                                                        B, Tmax, U, V, blank, normalized, lat2, den);
    return cudaGetLastError();
}

cudaError_t launch_logits_to_lattice(const void* logits, int dtype, const int* targets, const int* ll, const int* tl,
                                     int B, int Tmax, int U, int V, int blank, int normalized, float2* lat2, float* den,
                                     cudaStream_t st) {
    switch (dtype) {
        case 0: return launch_logits_to_lattice_t<float>(logits, targets, ll, tl, B, Tmax, U, V, blank, normalized, lat2, den, st);
        case 1: return launch_logits_to_lattice_t<__half>(logits, targets, ll, tl, B, Tmax, U, V, blank, normalized, lat2, den, st);
        case 2: return launch_logits_to_lattice_t<__nv_bfloat16>(logits, targets, ll, tl, B, Tmax, U, V, blank, normalized, lat2, den, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_alpha_beta(const float2* lat2, const int* ll, const int* tl, int B, int Tmax, int U, float* alpha,
                              float* beta, float* ll_alpha, float* ll_beta, float* cost, cudaStream_t st) {
    if (U > 1024) {  // wide lattice: several columns per thread (kMaxWideU bounds the shared-memory diagonal buffers)
        const size_t diag_bytes = 2 * (size_t)U * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(alpha_beta_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_bytes);
        if (e != cudaSuccess) return e;
        e = launch_pdl(alpha_beta_wide_kernel, dim3(B, 2), dim3(1024), diag_bytes, st, lat2, ll, tl, Tmax, U, alpha, beta, ll_alpha, ll_beta);
        if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return e;
        e = launch_pdl(finalize_cost_kernel, dim3((B + 127) / 128), dim3(128), 0, st, (const float*)ll_beta, cost, B);
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    const int threads = ((U + 31) / 32) * 32;
    const size_t ring_bytes = (size_t)kDpPrefetch * threads * sizeof(float2);
    // per launch, like the GEMM launchers: the attribute is per DEVICE, and one process may drive several GPUs
    cudaError_t e = cudaFuncSetAttribute(alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kDpPrefetch * 1024 * (int)sizeof(float2));
    if (e != cudaSuccess) return e;
    e = launch_pdl(alpha_beta_kernel, dim3(B, 2), dim3(threads), ring_bytes, st, lat2, ll, tl, Tmax, U, alpha, beta,
                               ll_alpha, ll_beta);
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return e;
    e = launch_pdl(finalize_cost_kernel, dim3((B + 127) / 128), dim3(128), 0, st, (const float*)ll_beta, cost, B);
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <typename T>
static cudaError_t launch_logits_grad_t(const void* logits, const int* targets, const int* ll, const int* tl, int B,
                                        int Tmax, int U, int V, int blank, const float2* lat2, const float* den,
                                        const float* alpha, const float* beta, const float* cost, const float* dcost,
                                        float clamp, void* dlogits, cudaStream_t st) {
    const long long cells = (long long)B * Tmax * U;
    const int wpb = 8;
    const long long blocks = (cells + wpb - 1) / wpb;
    logits_grad_kernel<T><<<(unsigned)blocks, wpb * 32, 0, st>>>(static_cast<const T*>(logits), targets, ll, tl, B, Tmax,
                                                                  U, V, blank, lat2, den, alpha, beta, cost, dcost,
                                                                  clamp, static_cast<T*>(dlogits));
    return cudaGetLastError();
}

cudaError_t launch_logits_grad(const void* logits, int dtype, const int* targets, const int* ll, const int* tl, int B,
                               int Tmax, int U, int V, int blank, const float2* lat2, const float* den,
                               const float* alpha, const float* beta, const float* cost, const float* dcost,
                               float clamp, void* dlogits, cudaStream_t st) {
    switch (dtype) {
        case 0: return launch_logits_grad_t<float>(logits, targets, ll, tl, B, Tmax, U, V, blank, lat2, den, alpha, beta, cost, dcost, clamp, dlogits, st);
        case 1: return launch_logits_grad_t<__half>(logits, targets, ll, tl, B, Tmax, U, V, blank, lat2, den, alpha, beta, cost, dcost, clamp, dlogits, st);
        case 2: return launch_logits_grad_t<__nv_bfloat16>(logits, targets, ll, tl, B, Tmax, U, V, blank, lat2, den, alpha, beta, cost, dcost, clamp, dlogits, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_logprobs_grad(const int* targets, const int* ll, const int* tl, int B, int Tmax, int U, int V,
                                 int blank, const float2* lat2, const float* alpha, const float* beta,
                                 const float* cost, const float* dcost, float* grads, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)B * Tmax * U * V, st);
    if (e != cudaSuccess) return e;
    const long long cells = (long long)B * Tmax * U;
    logprobs_grad_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(targets, ll, tl, B, Tmax, U, V, blank, lat2,
                                                                          alpha, beta, cost, dcost, grads);
    return cudaGetLastError();
}

}  // namespace tsasr
