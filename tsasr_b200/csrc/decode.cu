// decode.cu -- decode-time joint step: log_softmax(W * act(enc_t + dec) + bias) for B hypotheses.
//
// Replaces TransducerBeamSearcher._joint_forward_step (SB/decoders/transducer.py:375-384): Transducer_joint on
// [B,1,1,H] inputs (transducer_joint.py:73-74,95), the classifier Linear (linear.py:74) and LogSoftmax -- five tiny
// launches per decoded frame in the reference -- with two back-to-back launches of a few microseconds.  The step is
// a skinny GEMV batch (B <= a few dozen rows, W = V x H fp32 read once from L2), so it runs on the CUDA cores in fp32
// exactly like the eager path (same operands, no bf16 rounding):
//   joint_decode_logits_kernel : every CTA computes 8 vocabulary rows (one per warp, the W row held in registers)
//                                for all B hypotheses against the activations J kept in shared memory, writes the
//                                raw logits and its (max, sum-exp) partial per hypothesis;
//   joint_decode_norm_kernel   : one CTA per hypothesis folds the partials into log Z and normalises its row.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace tsasr {

static constexpr int kDecThreads = 256;
static constexpr int kDecRowsPerCta = 8;   // one vocabulary row per warp
static constexpr int kDecMaxB = 32;        // hypotheses per launch chunk (J tile in shared memory)
static constexpr int kDecMaxH4 = 12;       // float4 per lane of one W row: H <= 1536

__global__ void __launch_bounds__(kDecThreads)
joint_decode_logits_kernel(const float* __restrict__ enc, const float* __restrict__ dec, long long enc_stride,
                           long long dec_stride, const float* __restrict__ W, const float* __restrict__ bias, int B, int H,
                           int V, int act_kind, float act_param, float* __restrict__ out, float2* __restrict__ partials) {
    extern __shared__ __align__(16) float js[];           // [B][H] activations
    __shared__ float lg[kDecRowsPerCta][kDecMaxB];        // this CTA's logits
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = blockIdx.x * kDecRowsPerCta + warp;
    const int H4 = H >> 2;
    // the warp's W row: issued before the activations are built so that its L2 latency overlaps them
    float4 w[kDecMaxH4];
#pragma unroll
    for (int k = 0; k < kDecMaxH4; ++k) {
        const int i = lane + 32 * k;
        w[k] = (v < V && i < H4) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)v * H) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // rows may be strided views (tn_output[:, t, :]) or broadcast (stride 0): no host-side copies
    for (int i = threadIdx.x; i < B * H; i += kDecThreads) {
        const int b = i / H, h = i - b * H;
        js[i] = act_apply(enc[(long long)b * enc_stride + h] + dec[(long long)b * dec_stride + h], act_kind, act_param);
    }
    __syncthreads();

    if (v < V) {
        const float bv = bias ? bias[v] : 0.f;
        for (int b0 = 0; b0 < B; b0 += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < kDecMaxH4; ++k) {
                const int i = lane + 32 * k;
                if (i < H4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (b0 + j < B) {
                            const float4 x = reinterpret_cast<const float4*>(js + (size_t)(b0 + j) * H)[i];
                            acc[j] = fmaf(w[k].x, x.x, acc[j]); acc[j] = fmaf(w[k].y, x.y, acc[j]);
                            acc[j] = fmaf(w[k].z, x.z, acc[j]); acc[j] = fmaf(w[k].w, x.w, acc[j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
            }
            if (lane < 4 && b0 + lane < B) {
                const float y = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + bv;
                out[(size_t)(b0 + lane) * V + v] = y;
                lg[warp][b0 + lane] = y;
            }
        }
    }
    __syncthreads();
    // (max, sum-exp) of this CTA's rows, one hypothesis per thread
    const int nv = min(kDecRowsPerCta, V - blockIdx.x * kDecRowsPerCta);
    if (threadIdx.x < B) {
        const int b = threadIdx.x;
        float m = -INFINITY;
        for (int r = 0; r < nv; ++r) m = fmaxf(m, lg[r][b]);
        float s = 0.f;
        for (int r = 0; r < nv; ++r) s += __expf(lg[r][b] - m);
        partials[(size_t)b * gridDim.x + blockIdx.x] = make_float2(m, s);
    }
}

__global__ void __launch_bounds__(kDecThreads)
joint_decode_norm_kernel(float* __restrict__ out, const float2* __restrict__ partials, int V, int n_cta) {
    __shared__ float2 red[kDecThreads / 32];
    __shared__ float logz;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float m = -INFINITY, s = 0.f;
    for (int i = threadIdx.x; i < n_cta; i += kDecThreads) {
        const float2 pr = partials[(size_t)b * n_cta + i];
        const float mn = fmaxf(m, pr.x);
        s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + pr.y * __expf(pr.x - mn);
        m = mn;
    }
    auto combine = [](float& m, float& s, float m2, float s2) {
        const float mn = fmaxf(m, m2);
        if (mn == -INFINITY) return;  // both empty
        s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
        m = mn;
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) combine(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
    if (lane == 0) red[warp] = make_float2(m, s);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < kDecThreads / 32; ++k) combine(m, s, red[k].x, red[k].y);
        logz = m + __logf(s);
    }
    __syncthreads();
    const float lz = logz;
    float* row = out + (size_t)b * V;
    for (int i = threadIdx.x; i < V; i += kDecThreads) row[i] -= lz;
}

cudaError_t launch_joint_decode_step(const float* enc, const float* dec, long long enc_stride, long long dec_stride,
                                     const float* W, const float* bias, int B, int H, int V, int act_kind, float act_param,
                                     float* out, void* workspace, cudaStream_t st) {
    const int n_cta = (V + kDecRowsPerCta - 1) / kDecRowsPerCta;
    float2* partials = reinterpret_cast<float2*>(workspace);
    for (int b0 = 0; b0 < B; b0 += kDecMaxB) {
        const int nb = B - b0 < kDecMaxB ? B - b0 : kDecMaxB;
        const size_t smem = (size_t)nb * H * sizeof(float);
        cudaError_t e = cudaSuccess;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(joint_decode_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        joint_decode_logits_kernel<<<n_cta, kDecThreads, smem, st>>>(enc + (long long)b0 * enc_stride, dec + (long long)b0 * dec_stride,
                                                                     enc_stride, dec_stride, W, bias, nb, H, V, act_kind,
                                                                     act_param, out + (size_t)b0 * V, partials);
        joint_decode_norm_kernel<<<nb, kDecThreads, 0, st>>>(out + (size_t)b0 * V, partials, V, n_cta);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

size_t joint_decode_workspace_bytes(int V) {
    const int n_cta = (V + kDecRowsPerCta - 1) / kDecRowsPerCta;
    return sizeof(float2) * (size_t)kDecMaxB * n_cta;
}

}  // namespace tsasr
