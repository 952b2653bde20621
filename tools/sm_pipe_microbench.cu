// Issue-rate microbenchmark of the SM pipes the joint GEMM epilogue leans on (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sm_pipe_microbench tools/sm_pipe_microbench.cu
// Prints lane-operations per clock per SM for MUFU.EX2, FMNMX, FADD2 / FFMA2 (packed f32x2) and scalar FFMA at
// 2, 4, 8 and 16 resident warps per SM (one block per SM).
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void pipe_kernel(float* out, long long* cycles, int iters) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.001f * (threadIdx.x + i);
    float2 y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = make_float2(x[2 * i], x[2 * i + 1]);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[(i + 1) & 7]));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(x[(i + 1) & 7]));
            if (OP == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
        if (OP == 3 || OP == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned long long a = ((unsigned long long)__float_as_uint(y[i].y) << 32) | __float_as_uint(y[i].x);
                unsigned long long b = ((unsigned long long)__float_as_uint(y[(i + 1) & 3].y) << 32) | __float_as_uint(y[(i + 1) & 3].x);
                if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
                if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a) : "l"(b));
                y[i].x = __uint_as_float((unsigned)a);
                y[i].y = __uint_as_float((unsigned)(a >> 32));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += y[i].x + y[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int lane_ops_per_iter, float* out, long long* cyc) {
    const int iters = 4096;
    for (int warps : {1, 2, 4, 8, 16}) {
        pipe_kernel<OP><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        pipe_kernel<OP><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        printf("%-10s warps/SM=%2d  %.1f lane-ops/clk/SM\n", name, warps, (double)iters * lane_ops_per_iter * warps * 32 / avg);
    }
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    run<0>("MUFU.EX2", 8, out, cyc);
    run<5>("MUFU.LG2", 8, out, cyc);
    run<1>("FMNMX", 8, out, cyc);
    run<2>("FFMA", 8, out, cyc);
    run<3>("FADD2", 8, out, cyc);   // 4 packed instructions = 8 lane-ops per thread
    run<4>("FFMA2", 8, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
