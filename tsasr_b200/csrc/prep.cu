// prep.cu -- small launch-count savers on the host path into the fused loss (one launch instead of ~14).
//
//   lengths_kernel : relative -> absolute length conversion of SB/nnet/losses.py:58-59, bit-exact
//                    ((rel * dim) in fp32, round-half-to-even, int32), for both length vectors at once, plus the
//                    four statistics (max/min of each) that the torchaudio-style argument checks need.
//   cast3_kernel   : fp32 -> bf16 rounding of the three GEMM operands (enc_out, dec_out, W) in one launch.
//   prepare_inputs_kernel : both of the above plus the int64 -> int32 copy of the targets, in ONE launch
//                    (tsasr_joint_loss_fwd: the whole forward of the fused loss behind a single C call).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsasr {

__global__ void __launch_bounds__(256)
lengths_kernel(const float* __restrict__ rel_ll, const float* __restrict__ rel_tl, const int* __restrict__ abs_ll,
               const int* __restrict__ abs_tl, int B, int T, int n_targets, int* __restrict__ out_ll,
               int* __restrict__ out_tl, int* __restrict__ stats /* max_ll, max_tl, min_ll, min_tl */) {
    __shared__ int red[4][8];
    int mx_l = INT_MIN, mx_t = INT_MIN, mn_l = INT_MAX, mn_t = INT_MAX;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        // losses.py:58-59: (input_lens * logits.shape[1]).round().int(): fp32 product, rint, exact cast
        const int ll = rel_ll ? __float2int_rn(__fmul_rn(rel_ll[b], (float)T)) : abs_ll[b];
        const int tl = rel_tl ? __float2int_rn(__fmul_rn(rel_tl[b], (float)n_targets)) : abs_tl[b];
        if (out_ll) out_ll[b] = ll;
        if (out_tl) out_tl[b] = tl;
        mx_l = max(mx_l, ll); mn_l = min(mn_l, ll);
        mx_t = max(mx_t, tl); mn_t = min(mn_t, tl);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx_l = max(mx_l, __shfl_xor_sync(0xffffffffu, mx_l, o));
        mx_t = max(mx_t, __shfl_xor_sync(0xffffffffu, mx_t, o));
        mn_l = min(mn_l, __shfl_xor_sync(0xffffffffu, mn_l, o));
        mn_t = min(mn_t, __shfl_xor_sync(0xffffffffu, mn_t, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = mx_l; red[1][warp] = mx_t; red[2][warp] = mn_l; red[3][warp] = mn_t; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            mx_l = max(mx_l, red[0][w]); mx_t = max(mx_t, red[1][w]);
            mn_l = min(mn_l, red[2][w]); mn_t = min(mn_t, red[3][w]);
        }
        stats[0] = mx_l; stats[1] = mx_t; stats[2] = mn_l; stats[3] = mn_t;
    }
}

__global__ void __launch_bounds__(256)
cast3_kernel(const float4* __restrict__ a, size_t na4, const float4* __restrict__ b, size_t nb4,
             const float4* __restrict__ c, size_t nc4, uint2* __restrict__ oa, uint2* __restrict__ ob, uint2* __restrict__ oc) {
    const size_t total = na4 + nb4 + nc4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float4* src;
        uint2* dst;
        size_t k = i;
        if (k < na4) { src = a; dst = oa; }
        else if ((k -= na4) < nb4) { src = b; dst = ob; }
        else { k -= nb4; src = c; dst = oc; }
        const float4 v = src[k];
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        dst[k] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
}

// Everything the fused forward needs from its raw inputs in ONE launch: the three bf16 operand copies (skipped when the
// operands arrive as bf16), the int64 -> int32 copy of the targets (the recipe's tokens are int64; torchaudio's call site
// casts them, SB/nnet/losses.py:74), and -- block 0 -- the length conversion + statistics of lengths_kernel.
__global__ void __launch_bounds__(256)
prepare_inputs_kernel(const float4* __restrict__ a, size_t na4, const float4* __restrict__ b, size_t nb4,
                      const float4* __restrict__ c, size_t nc4, uint2* __restrict__ oa, uint2* __restrict__ ob, uint2* __restrict__ oc,
                      const long long* __restrict__ targets64, size_t n_targets_total, int* __restrict__ targets32,
                      const float* __restrict__ rel_ll, const float* __restrict__ rel_tl, const int* __restrict__ abs_ll,
                      const int* __restrict__ abs_tl, int B, int T, int n_targets, int* __restrict__ out_ll,
                      int* __restrict__ out_tl, int* __restrict__ stats, volatile int* __restrict__ stats_host, int stats_seq) {
    const size_t n_cast = na4 + nb4 + nc4;
    const size_t total = n_cast + (targets64 ? n_targets_total : 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (i >= n_cast) {
            targets32[i - n_cast] = (int)targets64[i - n_cast];
            continue;
        }
        const float4* src;
        uint2* dst;
        size_t k = i;
        if (k < na4) { src = a; dst = oa; }
        else if ((k -= na4) < nb4) { src = b; dst = ob; }
        else { k -= nb4; src = c; dst = oc; }
        const float4 v = src[k];
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        dst[k] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
    if (blockIdx.x != 0) return;
    __shared__ int red[4][8];
    int mx_l = INT_MIN, mx_t = INT_MIN, mn_l = INT_MAX, mn_t = INT_MAX;
    for (int bi = threadIdx.x; bi < B; bi += blockDim.x) {
        const int ll = rel_ll ? __float2int_rn(__fmul_rn(rel_ll[bi], (float)T)) : abs_ll[bi];   // losses.py:58-59, bit-exact
        const int tl = rel_tl ? __float2int_rn(__fmul_rn(rel_tl[bi], (float)n_targets)) : abs_tl[bi];
        out_ll[bi] = ll;
        out_tl[bi] = tl;
        mx_l = max(mx_l, ll); mn_l = min(mn_l, ll);
        mx_t = max(mx_t, tl); mn_t = min(mn_t, tl);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx_l = max(mx_l, __shfl_xor_sync(0xffffffffu, mx_l, o));
        mx_t = max(mx_t, __shfl_xor_sync(0xffffffffu, mx_t, o));
        mn_l = min(mn_l, __shfl_xor_sync(0xffffffffu, mn_l, o));
        mn_t = min(mn_t, __shfl_xor_sync(0xffffffffu, mn_t, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = mx_l; red[1][warp] = mx_t; red[2][warp] = mn_l; red[3][warp] = mn_t; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            mx_l = max(mx_l, red[0][w]); mx_t = max(mx_t, red[1][w]);
            mn_l = min(mn_l, red[2][w]); mn_t = min(mn_t, red[3][w]);
        }
        stats[0] = mx_l; stats[1] = mx_t; stats[2] = mn_l; stats[3] = mn_t;
        if (stats_host) {
            // zero-copy hand-off to the host (mapped pinned memory): the four numbers, a system-scope fence, then the
            // sequence tag the host polls for -- the argument checks need no stream, event or copy of their own
            stats_host[0] = mx_l; stats_host[1] = mx_t; stats_host[2] = mn_l; stats_host[3] = mn_t;
            __threadfence_system();
            stats_host[4] = stats_seq;
        }
    }
}

cudaError_t launch_prepare_inputs(const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc, void* oa, void* ob,
                                  void* oc, const long long* targets64, size_t n_targets_total, int* targets32, const float* rel_ll,
                                  const float* rel_tl, const int* abs_ll, const int* abs_tl, int B, int T, int n_targets, int* out_ll,
                                  int* out_tl, int* stats, int* stats_host, int stats_seq, int num_sms, cudaStream_t st) {
    const size_t total = (na + nb + nc) / 4 + (targets64 ? n_targets_total : 0);
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)num_sms * 16) blocks = (size_t)num_sms * 16;
    if (blocks == 0) blocks = 1;  // block 0 always runs: it converts the lengths
    prepare_inputs_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        reinterpret_cast<const float4*>(a), na / 4, reinterpret_cast<const float4*>(b), nb / 4, reinterpret_cast<const float4*>(c), nc / 4,
        reinterpret_cast<uint2*>(oa), reinterpret_cast<uint2*>(ob), reinterpret_cast<uint2*>(oc), targets64, n_targets_total, targets32,
        rel_ll, rel_tl, abs_ll, abs_tl, B, T, n_targets, out_ll, out_tl, stats, stats_host, stats_seq);
    return cudaGetLastError();
}

cudaError_t launch_lengths(const float* rel_ll, const float* rel_tl, const int* abs_ll, const int* abs_tl, int B, int T,
                           int n_targets, int* out_ll, int* out_tl, int* stats, cudaStream_t st) {
    lengths_kernel<<<1, 256, 0, st>>>(rel_ll, rel_tl, abs_ll, abs_tl, B, T, n_targets, out_ll, out_tl, stats);
    return cudaGetLastError();
}

cudaError_t launch_cast3(const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc, void* oa, void* ob,
                         void* oc, int num_sms, cudaStream_t st) {
    const size_t total4 = (na + nb + nc) / 4;
    size_t blocks = (total4 + 255) / 256;
    if (blocks > (size_t)num_sms * 16) blocks = (size_t)num_sms * 16;
    if (blocks == 0) return cudaSuccess;
    cast3_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(a), na / 4, reinterpret_cast<const float4*>(b),
                                                   nb / 4, reinterpret_cast<const float4*>(c), nc / 4,
                                                   reinterpret_cast<uint2*>(oa), reinterpret_cast<uint2*>(ob), reinterpret_cast<uint2*>(oc));
    return cudaGetLastError();
}

}  // namespace tsasr
