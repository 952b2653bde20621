// Micro-benchmarks of the tcgen05 issue path (development tool): how long do back-to-back MMAs, commits
// and mbarrier try_waits take for the single issuing lane?  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cstdio>
#include "../tsasr_b200/csrc/common.cuh"
using namespace tsasr;

__global__ void __launch_bounds__(128, 1) k(long long* out, int n_mma, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tp;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00;
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc<512>(&tp);
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tb = tp;
    if (threadIdx.x == 0) {
        const uint64_t a = make_smem_desc_sw128(smem_u32(smem), 0, 1024);
        const uint64_t b = make_smem_desc_sw128(smem_u32(smem) + 16384, 0, 1024);
        const uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
        long long t0 = clock64();
        if (mode == 0) {            // back-to-back MMAs, one commit at the end, then wait
            for (int i = 0; i < n_mma; ++i) umma_bf16(tb, a, b, idesc, 1);
            long long t1 = clock64();
            umma_commit(&bar[0]);
            mbar_wait(&bar[0], 0);
            long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        } else if (mode == 1) {     // 2 MMAs + commit per "stage" (single-CTA kernel pattern), no waits
            for (int i = 0; i < n_mma / 2; ++i) { umma_bf16(tb, a, b, idesc, 1); umma_bf16(tb, a, b, idesc, 1); umma_commit(&bar[1]); }
            long long t1 = clock64();
            umma_commit(&bar[0]); mbar_wait(&bar[0], 0);
            long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        } else if (mode == 2) {     // try_wait on an already-completed barrier, n times
            mbar_arrive(&bar[2]);
            uint32_t acc = 0;
            for (int i = 0; i < n_mma; ++i) acc += mbar_try_wait(&bar[2], 0);
            long long t1 = clock64();
            out[0] = t1 - t0; out[1] = acc;
        } else if (mode == 3) {     // 4 MMAs + commit, then wait for that commit (fully serialised stage)
            uint32_t ph = 0;
            for (int i = 0; i < n_mma / 4; ++i) {
                for (int kk = 0; kk < 4; ++kk) umma_bf16(tb, a, b, idesc, 1);
                umma_commit(&bar[3]); mbar_wait(&bar[3], ph); ph ^= 1;
            }
            long long t1 = clock64();
            out[0] = t1 - t0; out[1] = 0;
        }
    }
    tcgen05_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[] = {"back-to-back MMAs (M128 N256 K16)", "2 MMAs + commit per stage", "try_wait on completed barrier", "4 MMAs + commit + wait (serialised)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int n : {64, 512}) {
            for (int rep = 0; rep < 2; ++rep) k<<<1, 128, 64 * 1024>>>(d, n, mode);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-40s n=%4d issue=%8lld cyc (%.1f/op) total=%8lld cyc (%.1f/op) %s\n", names[mode], n, h[0], (double)h[0] / n, h[1], (double)h[1] / n,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
