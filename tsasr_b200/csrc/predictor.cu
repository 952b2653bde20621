// predictor.cu -- the prediction network of the transducer: one-hot embedding -> 1-layer LSTM.
//
// Reference: speechbrain.nnet.embedding.Embedding with consider_as_one_hot=True (SB/nnet/embedding.py:65-114: a frozen
// eye matrix with the blank row zeroed, so the "embedding" of token k is the one-hot vector e_{k - [k > blank]} of size V-1,
// zeros for the blank) followed by speechbrain.nnet.RNN.LSTM (SB/nnet/RNN.py:170-278: torch.nn.LSTM, batch_first, fed a
// PackedSequence built from RELATIVE lengths, :25-38,262-276), as chained at train_librispeechmix_scratch.py:125-126 and
// configured at hparams/LibriSpeechMix/conformer-t_scratch.yaml:176-185.  The reference materialises the [B,U,V-1] one-hot
// tensor, multiplies it with W_ih (a 6.5 GFLOP fp32 GEMM whose rows are 99.9 % zeros), copies the lengths to the host
// (`.cpu()`, a stream synchronisation every step) and runs cuDNN's LSTM; its backward repeats the dense GEMM for dW_ih.
//
// Here:
//   lstm_seq_fwd_kernel  : the whole teacher-forced recurrence in ONE cooperative launch.  CTA c owns 4 hidden units (its 16
//                          rows of W_hh live in registers, 64 per thread); a step is 16 x Hd dot products per batch tile on the
//                          CUDA cores in fp32 (the reference's arithmetic: no bf16 anywhere, the state feeds back 100 times),
//                          a 31-shuffle butterfly, the cell update, and one grid-wide hand-off of h_t through global memory
//                          (carried by the data itself: sentinel-filled buffer, polled with cp.async, see below).  x_t W_ih^T is a COLUMN GATHER of W_ih by token id --
//                          no one-hot tensor, no GEMM -- prefetched a step ahead.  The relative -> absolute length conversion
//                          (fp32 product, truncation: what torch's pack_padded_sequence does with the float lengths it is
//                          given) happens on the device: no host synchronisation.
//   lstm_seq_bwd_kernel  : back-propagation through time in ONE cooperative launch with the same ownership (4 units per CTA,
//                          the 4 x 4Hd slice of W_hh^T in registers); emits dG = d loss / d gate pre-activations [B,U,4Hd].
//   onehot_dw_kernel     : dW_ih[:, v] = sum of dG rows whose token maps to column v, in a fixed order (deterministic; the
//                          reference's dense GEMM against the one-hot tensor collapses to this gather-sum).
// dW_hh = dG^T h_prev and db = column sums of dG are one tsasr_linear_bwd call (linear.cuh), made by the host side.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "predictor.cuh"

namespace tsasr {

// ------------------------------------------------------------------------------------------------------------------
// Grid-wide hand-off of one recurrence step, carried BY THE DATA.  The exchange buffer (h_t in the forward, the gate
// gradients in the backward) is pre-filled with a sentinel bit pattern (0xFFFFFFFF: a NaN no arithmetic produces) by a
// cudaMemsetAsync in front of the launch; every element is written exactly once, by the thread that owns it, at its
// step.  A consumer copies the 16-utterance tile of the previous step into shared memory with cp.async (.cg: L2 only)
// and re-fetches the 16-byte pieces that still show the sentinel until none does.  There is no flag, no fence and no
// atomic on the critical path of a step: producer store -> L2 -> consumer copy.  Measured for 100 steps of the forward
// (B = 16, Hd = 512): arrival counter + __threadfence 420 us, one flag per CTA polled by a warp 643 us, this scheme see
// profiles/r2_predictor.txt.  All CTAs are co-resident (cooperative launch), so the polling cannot deadlock, and a step
// can run at most one step ahead of the slowest CTA, so no buffer is reused.  Fail-stop after kMbarTimeoutNs like mbar_wait.
// ------------------------------------------------------------------------------------------------------------------
static constexpr unsigned int kSentinelBits = 0xFFFFFFFFu;
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ bool has_sentinel(const float4& v) {
    return (__float_as_uint(v.x) == kSentinelBits) | (__float_as_uint(v.y) == kSentinelBits) | (__float_as_uint(v.z) == kSentinelBits) |
           (__float_as_uint(v.w) == kSentinelBits);
}
// tile row bb (0..15) = utterance b0 + bb: ROW_F4 float4 at gbase + (b0 + bb) * row_stride (floats); rows past B are zeros
// `probe`: 16 bytes per producer CTA (its 4 units of the LAST row of the tile, written among the last of its step), polled
// first by one thread per producer: the bulk copy then starts when the step's results are (almost certainly) all there,
// instead of fetching a tile full of sentinels and re-fetching most of it -- in the backward a tile is 128 KB per CTA and
// a wasted round costs more than the step's arithmetic (measured: fetch 5.5 -> see profiles/r2_predictor.txt us per step).
// Correctness never depends on the probe: every piece is still checked for the sentinel after the copy.
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
template <int ROW_F4>
__device__ __forceinline__ void fetch_tile_polling(float* smem_dst, const float* gbase, size_t row_stride, int b0, int B, uint32_t tag,
                                                   const float* probe) {
    constexpr int kTotal = kLstmBatchTile * ROW_F4;
    if (probe != nullptr && threadIdx.x < gridDim.x) {
        const float* pp = probe + 4 * threadIdx.x;
        uint32_t spins = 0;
        unsigned long long t0 = 0;
        while (has_sentinel(ld_volatile_f4(pp))) {
#if !defined(TSASR_NO_WATCHDOG)
            if ((++spins & 4095u) == 0u) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kMbarTimeoutNs) {
                    g_tsasr_hang_info[0] = 0xDEAD0000u | tag | 0x80u;
                    g_tsasr_hang_info[1] = blockIdx.x;
                    g_tsasr_hang_info[2] = threadIdx.x;
                    g_tsasr_hang_info[3] = (unsigned)b0;
                    __threadfence_system();
                    __trap();
                }
            }
#endif
        }
    }
    if (probe != nullptr) __syncthreads();
    for (int i = threadIdx.x; i < kTotal; i += kLstmThreads) {
        const int bb = i / ROW_F4, k4 = i - bb * ROW_F4;
        if (b0 + bb < B) cp_async_16(smem_dst + 4 * i, gbase + (size_t)(b0 + bb) * row_stride + 4 * k4);
        else reinterpret_cast<float4*>(smem_dst)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait_all();
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        bool pending = false;
        for (int i = threadIdx.x; i < kTotal; i += kLstmThreads) {
            const int bb = i / ROW_F4, k4 = i - bb * ROW_F4;
            if (b0 + bb < B && has_sentinel(reinterpret_cast<const float4*>(smem_dst)[i])) {
                cp_async_16(smem_dst + 4 * i, gbase + (size_t)(b0 + bb) * row_stride + 4 * k4);
                pending = true;
            }
        }
        if (!pending) break;
        cp_async_wait_all();
#if !defined(TSASR_NO_WATCHDOG)
        if ((++spins & 1023u) == 0u) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kMbarTimeoutNs) {
                g_tsasr_hang_info[0] = 0xDEAD0000u | tag;
                g_tsasr_hang_info[1] = blockIdx.x;
                g_tsasr_hang_info[2] = threadIdx.x;
                g_tsasr_hang_info[3] = (unsigned)b0;
                __threadfence_system();
                __trap();
            }
        }
#endif
    }
    __syncthreads();
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// sum over the 32 lanes of v[i] for every i; afterwards lane l holds the total of v[l] in v[0]
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            v[i] = (up ? v[i + off] : v[i]) + recv;
        }
    }
    return v[0];
}

// column of W_ih selected by a token (SB/nnet/embedding.py:88-100): -1 for the blank (all-zero embedding)
__device__ __forceinline__ int onehot_column(long long tok, int blank, int n_embed) {
    if (tok == blank) return -1;
    const long long c = tok > blank ? tok - 1 : tok;
    return (c >= 0 && c < n_embed) ? (int)c : -1;
}

template <int KPL>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_seq_fwd_kernel(const LstmFwdParams p) {
    constexpr int Hd = KPL * 32;
    extern __shared__ __align__(16) float dyn_smem[];
    float* h_s = dyn_smem;                                            // [16][Hd] state of the previous step, one batch tile
    int* col_s = reinterpret_cast<int*>(dyn_smem + kLstmBatchTile * Hd);  // [B*U] W_ih column per position (one-hot mode)
    __shared__ int len_s[kLstmBatchTile * kLstmMaxPasses];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jj = warp & 3, bg = warp >> 2;
    const int j = blockIdx.x * kLstmUnits + jj;
    const int B = p.B, U = p.U;
    const int passes = (B + kLstmBatchTile - 1) / kLstmBatchTile;
    const bool onehot = p.xw == nullptr;

    // this thread's slice of W_hh: rows g*Hd + j, columns (q*128 + lane*4 .. +3): consecutive lanes read consecutive float4
    float w[4][KPL];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int q = 0; q < KPL / 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.W_hh + (size_t)(g * Hd + j) * Hd + q * 128 + lane * 4));
            w[g][4 * q] = v.x; w[g][4 * q + 1] = v.y; w[g][4 * q + 2] = v.z; w[g][4 * q + 3] = v.w;
        }
    for (int b = threadIdx.x; b < B; b += kLstmThreads) {
        // SB/nnet/RNN.py:35 + torch.nn.utils.rnn.pack_padded_sequence: (rel * U) in fp32, then an int64 cast (truncation)
        int L = p.rel_lengths ? __float2int_rz(__fmul_rn(p.rel_lengths[b], (float)U)) : p.abs_lengths[b];
        L = min(max(L, 0), U);
        len_s[b] = L;
        if (blockIdx.x == 0 && p.lengths_out) p.lengths_out[b] = L;
    }
    if (onehot) {
        for (int i = threadIdx.x; i < B * U; i += kLstmThreads)
            col_s[i] = onehot_column(p.tok64 ? p.tok64[i] : (long long)p.tok32[i], p.blank, p.n_embed);
    }
    // the cell threads: lanes 0-7 of every warp own (batch bg*8 + lane of each pass, unit j)
    float bias_g[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane < 8) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (p.b_hh) bias_g[g] += __ldg(p.b_hh + g * Hd + j);
            if (onehot && p.b_ih) bias_g[g] += __ldg(p.b_ih + g * Hd + j);  // dense mode: xw already holds x W_ih^T + b_ih
        }
    }
    float c_state[kLstmMaxPasses], h_state[kLstmMaxPasses];
#pragma unroll
    for (int ps = 0; ps < kLstmMaxPasses; ++ps) { c_state[ps] = 0.f; h_state[ps] = 0.f; }
    __syncthreads();

    for (int u = 0; u < U; ++u) {
        // input contribution of this step, issued before the hand-off wait so that its latency hides behind it
        float xin[kLstmMaxPasses][4];
#pragma unroll
        for (int ps = 0; ps < kLstmMaxPasses; ++ps) {
            const int b = ps * kLstmBatchTile + bg * 8 + lane;
#pragma unroll
            for (int g = 0; g < 4; ++g) xin[ps][g] = 0.f;
            if (lane < 8 && ps < passes && b < B) {
                if (onehot) {
                    const int col = col_s[b * U + u];
                    if (col >= 0) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) xin[ps][g] = __ldg(p.W_ih + (size_t)(g * Hd + j) * p.n_embed + col);
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) xin[ps][g] = __ldg(p.xw + ((size_t)b * U + u) * (4 * Hd) + g * Hd + j);
                }
            }
        }
#pragma unroll
        for (int ps = 0; ps < kLstmMaxPasses; ++ps) {
            if (ps >= passes) break;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
            if (u > 0) {
                // h_{u-1} of this batch tile: out[b, u-1, :], as soon as its owners have written it
                const int b_probe = min(ps * kLstmBatchTile + kLstmBatchTile - 1, B - 1);  // last utterance of this tile
                if (!(p.dbg & 2))
                    fetch_tile_polling<Hd / 4>(h_s, p.out + (size_t)(u - 1) * Hd, (size_t)U * Hd, ps * kLstmBatchTile, B, 0xA00,
                                               (p.dbg & 4) ? nullptr : p.out + ((size_t)b_probe * U + (u - 1)) * Hd);
                if (!(p.dbg & 1))
#pragma unroll
                for (int q = 0; q < KPL / 4; ++q) {
#pragma unroll
                    for (int bb = 0; bb < 8; ++bb) {
                        const float4 hv = *reinterpret_cast<const float4*>(h_s + (bg * 8 + bb) * Hd + q * 128 + lane * 4);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float a = v[g * 8 + bb];
                            a = fmaf(w[g][4 * q], hv.x, a);
                            a = fmaf(w[g][4 * q + 1], hv.y, a);
                            a = fmaf(w[g][4 * q + 2], hv.z, a);
                            a = fmaf(w[g][4 * q + 3], hv.w, a);
                            v[g * 8 + bb] = a;
                        }
                    }
                }
                butterfly32(v, lane);  // lane l: gate l >> 3, batch l & 7
            }
            const float tot = v[0];
            const int src = lane & 7;
            const float a_i = __shfl_sync(0xffffffffu, tot, src), a_f = __shfl_sync(0xffffffffu, tot, src + 8);
            const float a_g = __shfl_sync(0xffffffffu, tot, src + 16), a_o = __shfl_sync(0xffffffffu, tot, src + 24);
            const int b = ps * kLstmBatchTile + bg * 8 + lane;
            if (lane < 8 && b < B) {
                const int L = len_s[b];
                const bool valid = u < L;
                const float gi = sigmoid_f(a_i + xin[ps][0] + bias_g[0]), gf = sigmoid_f(a_f + xin[ps][1] + bias_g[1]);
                const float gg = tanhf(a_g + xin[ps][2] + bias_g[2]), go = sigmoid_f(a_o + xin[ps][3] + bias_g[3]);
                const float c = fmaf(gf, c_state[ps], gi * gg);
                const float h = go * tanhf(c);
                const size_t pos = (size_t)b * U + u;
                if (valid) {
                    c_state[ps] = c;
                    h_state[ps] = h;
                    if (p.gates) {
                        float* gp = p.gates + pos * (4 * Hd) + j;
                        gp[0] = gi; gp[Hd] = gf; gp[2 * Hd] = gg; gp[3 * Hd] = go;
                    }
                    if (p.cells) p.cells[pos * Hd + j] = c;
                }
                // padded positions read as zeros (pad_packed_sequence, SB/nnet/RNN.py:41-54); their state is frozen
                p.out[pos * Hd + j] = valid ? h : 0.f;
                if (p.hprev) {
                    if (u == 0) p.hprev[pos * Hd + j] = 0.f;
                    if (u + 1 < U) p.hprev[(pos + 1) * Hd + j] = valid ? h : 0.f;
                }
                if (u == U - 1) {  // final state of every utterance = state after its last valid step (packed semantics)
                    if (p.h_n) p.h_n[(size_t)b * Hd + j] = h_state[ps];
                    if (p.c_n) p.c_n[(size_t)b * Hd + j] = c_state[ps];
                }
            }
            __syncthreads();  // h_s is overwritten by the next batch tile / the next step
        }
    }
}

template <int KPL>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_seq_bwd_kernel(const LstmBwdParams p) {
    constexpr int Hd = KPL * 32, G = 4 * Hd;
    extern __shared__ __align__(16) float dyn_smem[];
    float* dg_s = dyn_smem;                       // [16][4Hd] gate gradients of step u+1, one batch tile
    float* red = dyn_smem + kLstmBatchTile * G;   // [4 gate quarters][2 batch groups][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = warp & 3, bg = warp >> 2;
    const int B = p.B, U = p.U;
    const int passes = (B + kLstmBatchTile - 1) / kLstmBatchTile;
    // this thread's slice of W_hh^T: rows g*Hd + k (k = q*128 + lane*4 + e), the CTA's 4 columns -> one float4 per row
    float w[4][KPL];
#pragma unroll
    for (int kk = 0; kk < KPL; ++kk) {
        const int k = (kk >> 2) * 128 + lane * 4 + (kk & 3);
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.W_hh + (size_t)(g * Hd + k) * Hd + blockIdx.x * kLstmUnits));
        w[0][kk] = v.x; w[1][kk] = v.y; w[2][kk] = v.z; w[3][kk] = v.w;
    }
    // cell threads: warps 0 and 1 (batch group = warp), lane l -> unit l >> 3, batch l & 7
    const bool cell_thread = warp < 2;
    const int cj = blockIdx.x * kLstmUnits + (lane >> 3);
    float dc_next[kLstmMaxPasses];
#pragma unroll
    for (int ps = 0; ps < kLstmMaxPasses; ++ps) dc_next[ps] = 0.f;

    for (int u = U - 1; u >= 0; --u) {
        // operands of this step's cell gradient, issued before the hand-off wait
        float sv[kLstmMaxPasses][7];  // i, f, g, o, c, c_prev, d_out
        int len_b[kLstmMaxPasses];
#pragma unroll
        for (int ps = 0; ps < kLstmMaxPasses; ++ps) {
            const int b = ps * kLstmBatchTile + warp * 8 + (lane & 7);
            len_b[ps] = 0;
#pragma unroll
            for (int i = 0; i < 7; ++i) sv[ps][i] = 0.f;
            if (cell_thread && ps < passes && b < B) {
                len_b[ps] = min(max(__ldg(p.lengths + b), 0), U);
                if (u < len_b[ps]) {
                    const size_t pos = (size_t)b * U + u;
                    const float* gp = p.gates + pos * G + cj;
                    sv[ps][0] = __ldg(gp); sv[ps][1] = __ldg(gp + Hd); sv[ps][2] = __ldg(gp + 2 * Hd); sv[ps][3] = __ldg(gp + 3 * Hd);
                    sv[ps][4] = __ldg(p.cells + pos * Hd + cj);
                    sv[ps][5] = u > 0 ? __ldg(p.cells + (pos - 1) * Hd + cj) : 0.f;
                    sv[ps][6] = __ldg(p.d_out + pos * Hd + cj);
                    if (u == len_b[ps] - 1 && p.d_hn) sv[ps][6] += __ldg(p.d_hn + (size_t)b * Hd + cj);
                }
            }
        }
#pragma unroll
        for (int ps = 0; ps < kLstmMaxPasses; ++ps) {
            if (ps >= passes) break;
            if (u < U - 1) {
                // gate gradients of step u+1 of this batch tile, as soon as their owners have written them
                const int b_probe = min(ps * kLstmBatchTile + kLstmBatchTile - 1, B - 1);  // last utterance of this tile
                if (!(p.dbg & 2))
                    fetch_tile_polling<G / 4>(dg_s, p.dG + (size_t)(u + 1) * G, (size_t)U * G, ps * kLstmBatchTile, B, 0xB00,
                                              (p.dbg & 4) ? nullptr : p.dG + ((size_t)b_probe * U + (u + 1)) * G + 3 * Hd);
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
                if (!(p.dbg & 1))
#pragma unroll
                for (int q = 0; q < KPL / 4; ++q) {
#pragma unroll
                    for (int bb = 0; bb < 8; ++bb) {
                        const float4 dv = *reinterpret_cast<const float4*>(dg_s + (bg * 8 + bb) * G + g * Hd + q * 128 + lane * 4);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            float a = v[jj * 8 + bb];
                            a = fmaf(w[jj][4 * q], dv.x, a);
                            a = fmaf(w[jj][4 * q + 1], dv.y, a);
                            a = fmaf(w[jj][4 * q + 2], dv.z, a);
                            a = fmaf(w[jj][4 * q + 3], dv.w, a);
                            v[jj * 8 + bb] = a;
                        }
                    }
                }
                butterfly32(v, lane);  // lane l: unit l >> 3, batch l & 7, this warp's gate quarter
                red[(g * 2 + bg) * 32 + lane] = v[0];
                __syncthreads();
            }
            const int b = ps * kLstmBatchTile + warp * 8 + (lane & 7);
            if (cell_thread && b < B) {
                const size_t pos = (size_t)b * U + u;
                float* dgp = p.dG + pos * G + cj;
                if (u < len_b[ps]) {
                    float dh = sv[ps][6];
                    if (u < U - 1) dh += (red[(0 * 2 + warp) * 32 + lane] + red[(1 * 2 + warp) * 32 + lane]) +
                                         (red[(2 * 2 + warp) * 32 + lane] + red[(3 * 2 + warp) * 32 + lane]);
                    const float gi = sv[ps][0], gf = sv[ps][1], gg = sv[ps][2], go = sv[ps][3], c = sv[ps][4], cp = sv[ps][5];
                    const float tc = tanhf(c);
                    float dc = fmaf(dh * go, 1.f - tc * tc, dc_next[ps]);
                    if (u == len_b[ps] - 1 && p.d_cn) dc += __ldg(p.d_cn + (size_t)b * Hd + cj);
                    dgp[0] = dc * gg * gi * (1.f - gi);
                    dgp[Hd] = dc * cp * gf * (1.f - gf);
                    dgp[2 * Hd] = dc * gi * (1.f - gg * gg);
                    dgp[3 * Hd] = dh * tc * go * (1.f - go);
                    dc_next[ps] = dc * gf;
                } else {  // padded position: contributes nothing, carries nothing
                    dgp[0] = 0.f; dgp[Hd] = 0.f; dgp[2 * Hd] = 0.f; dgp[3 * Hd] = 0.f;
                    dc_next[ps] = 0.f;
                }
            }
            __syncthreads();  // dg_s / red are overwritten by the next batch tile / the next step
        }
    }
}

// dW_ih[r, v] = sum over positions (b,u) whose token maps to column v of dG[b,u,r], positions in increasing order.
// One CTA per 8 consecutive columns.  Positions are scanned in chunks of kDwChunk: warp c builds the ordered match list of
// column v0 + c for the chunk (ballot compaction), then every thread adds the matching dG rows to its 8 rows x 8 columns of
// register accumulators (coalesced reads along r) and finally writes 8 consecutive floats per row.
static constexpr int kDwChunk = 1024;
__global__ void __launch_bounds__(256) onehot_dw_kernel(const long long* __restrict__ tok64, const int* __restrict__ tok32, int blank,
                                                        int n_embed, const float* __restrict__ dG, int n_pos, int G,
                                                        float* __restrict__ dW) {
    __shared__ int pos_s[8][kDwChunk];
    __shared__ int cnt_s[8];
    const int v0 = blockIdx.x * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r0 = 0; r0 < G; r0 += 256 * 8) {
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
        for (int base = 0; base < n_pos; base += kDwChunk) {
            const int v = v0 + warp, end = min(n_pos, base + kDwChunk);
            int n = 0;
            for (int i0 = base; i0 < end; i0 += 32) {
                const int i = i0 + lane;
                bool hit = false;
                if (i < end && v < n_embed) hit = onehot_column(tok64 ? tok64[i] : (long long)tok32[i], blank, n_embed) == v;
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) pos_s[warp][n + __popc(m & ((1u << lane) - 1u))] = i;
                n += __popc(m);
            }
            if (lane == 0) cnt_s[warp] = n;
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int nc = cnt_s[c];
                for (int k = 0; k < nc; ++k) {
                    const float* row = dG + (size_t)pos_s[c][k] * G + r0 + threadIdx.x;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (r0 + threadIdx.x + 256 * i < G) acc[i][c] += __ldg(row + 256 * i);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + threadIdx.x + 256 * i;
            if (r >= G) continue;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (v0 + c < n_embed) dW[(size_t)r * n_embed + v0 + c] = acc[i][c];
        }
    }
}

// ---- launchers ----
template <typename P>
static cudaError_t launch_coop(void (*kern)(const P), const P& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLstmThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm * sms < grid) return cudaErrorCooperativeLaunchTooLarge;  // all CTAs must be co-resident (grid-wide hand-off)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kLstmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

size_t lstm_fwd_smem_bytes(int B, int U, int Hd, bool onehot) {
    return (size_t)kLstmBatchTile * Hd * 4 + (onehot ? (size_t)B * U * 4 : 0);
}

cudaError_t launch_lstm_fwd(const LstmFwdParams& p, cudaStream_t st) {
    const int grid = p.Hd / kLstmUnits;
    const size_t smem = lstm_fwd_smem_bytes(p.B, p.U, p.Hd, p.xw == nullptr);
    switch (p.Hd) {
        case 128: return launch_coop(lstm_seq_fwd_kernel<4>, p, grid, smem, st);
        case 256: return launch_coop(lstm_seq_fwd_kernel<8>, p, grid, smem, st);
        case 512: return launch_coop(lstm_seq_fwd_kernel<16>, p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_lstm_bwd(const LstmBwdParams& p, cudaStream_t st) {
    const int grid = p.Hd / kLstmUnits;
    const size_t smem = (size_t)kLstmBatchTile * 4 * p.Hd * 4 + 4 * 2 * 32 * 4;
    switch (p.Hd) {
        case 128: return launch_coop(lstm_seq_bwd_kernel<4>, p, grid, smem, st);
        case 256: return launch_coop(lstm_seq_bwd_kernel<8>, p, grid, smem, st);
        case 512: return launch_coop(lstm_seq_bwd_kernel<16>, p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_onehot_dw(const void* tokens, int tokens_i64, int blank, int n_embed, const float* dG, int n_pos, int G, float* dW,
                             cudaStream_t st) {
    onehot_dw_kernel<<<(n_embed + 7) / 8, 256, 0, st>>>(tokens_i64 ? static_cast<const long long*>(tokens) : nullptr,
                                                            tokens_i64 ? nullptr : static_cast<const int*>(tokens), blank, n_embed, dG,
                                                            n_pos, G, dW);
    return cudaGetLastError();
}

}  // namespace tsasr
