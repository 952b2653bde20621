// Micro-benchmark (development tool): cycles per tcgen05.mma as a function of N, operand majors and cta_group.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/umma_shapes tools/umma_shapes.cu
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include "../tsasr_b200/csrc/common.cuh"
using namespace tsasr;

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) k(long long* out, int n_mma, int N, int a_mn, int b_mn, int n_alt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tp;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00;
    fence_proxy_async_smem();
    if (threadIdx.x < 32) { if (PAIR) tmem_alloc_2cta<512>(&tp); else tmem_alloc<512>(&tp); }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = tp;
    if (threadIdx.x < 32 && rank == 0) {
        const int M = PAIR ? 256 : 128;
        // K-major: rows of 128 B (64 k), 8-row atoms 1024 B apart; MN-major: 64-wide blocks 16 KB apart, 8-k groups 1024 B apart
        const uint64_t a = a_mn ? make_smem_desc_sw128(smem_u32(smem), 16384, 1024) : make_smem_desc_sw128(smem_u32(smem), 0, 1024);
        const uint64_t b = b_mn ? make_smem_desc_sw128(smem_u32(smem) + 32768, 16384, 1024) : make_smem_desc_sw128(smem_u32(smem) + 32768, 0, 1024);
        const uint32_t idesc = make_idesc_bf16(M, N, a_mn, b_mn);
        const uint32_t idesc2 = make_idesc_bf16(M, n_alt > 0 ? n_alt : N, a_mn, b_mn);
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            if (PAIR) umma_bf16_2cta_e(tb, a, b, (n_alt > 0 && (i & 1)) ? idesc2 : idesc, 1);
            else umma_bf16_e(tb, a, b, (n_alt > 0 && (i & 1)) ? idesc2 : idesc, 1);
        }
        long long t1 = clock64();
        if (PAIR) umma_commit_2cta_e(&bar[0], 1); else umma_commit_e(&bar[0]);
        mbar_wait(&bar[0], 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x < 32) { tcgen05_fence_after(); if (PAIR) tmem_dealloc_2cta<512>(tb); else tmem_dealloc<512>(tb); }
}

template <bool PAIR>
static void run(long long* d, int N, int a_mn, int b_mn, int n_alt) {
    const int n = 1024;
    auto kern = k<PAIR>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(PAIR ? 2 : 1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 160 * 1024;
    cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, kern, d, n, N, a_mn, b_mn, n_alt);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const int M = PAIR ? 256 : 128;
    const double flops = 2.0 * M * (n_alt > 0 ? 0.5 * (N + n_alt) : N) * 16 / (PAIR ? 2 : 1);
    printf("cta_group::%d M=%3d N=%3d%s A=%s B=%s : %7.1f cyc/mma  %6.0f flop/cyc/SM %s\n", PAIR ? 2 : 1, M, N,
           n_alt > 0 ? "/alt" : "    ", a_mn ? "MN" : "K ", b_mn ? "MN" : "K ", (double)h[1] / n, flops / ((double)h[1] / n),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    for (int maj = 0; maj < 4; ++maj)
        for (int N : {32, 64, 128, 256}) run<false>(d, N, maj >> 1, maj & 1, 0);
    for (int maj = 0; maj < 4; ++maj)
        for (int N : {32, 64, 128, 256}) run<true>(d, N, maj >> 1, maj & 1, 0);
    run<true>(d, 256, 1, 1, 32);   // alternating N=256 / N=32 (dW + db pattern)
    run<true>(d, 128, 1, 1, 32);
    return 0;
}
