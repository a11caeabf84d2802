// Microbenchmark: TMEM read bandwidth (tcgen05.ld 32x32b.x32) and MUFU.EX2 throughput per SM on sm_100a.
// One CTA per SM; W warps (W = 4, 8, 16) sweep a 512-column allocation R times.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mcm_b200/csrc/ptx.cuh"
using namespace mcm;

__global__ void tmem_read_kernel(int reps, int cols_per_warp_span, long long* cycles, float* sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&tptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int c = 0; c < cols_per_warp_span; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(base + ((c + (warp >> 2) * 64) & 511), v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 8) acc += __uint_as_float(v[e]);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tptr);
}

__global__ void ex2_kernel(int reps, long long* cycles, float* sink) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456f) sink[0] = s;
}

int main() {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 4);
    long long h[148];
    for (int warps : {4, 8, 16}) {
        const int reps = 200, span = 256;
        tmem_read_kernel<<<148, warps * 32>>>(reps, span, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double bytes = (double)warps * reps * (span / 32) * 32 * 32 * 4;
        printf("tmem_ld x32: %2d warps: %lld cycles, %.1f B/clk/SM (%s)\n", warps, h[0], bytes / h[0], cudaGetErrorString(e));
    }
    for (int warps : {4, 8, 16, 32}) {
        const int reps = 2000;
        ex2_kernel<<<148, warps * 32>>>(reps, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double ops = (double)warps * 32 * reps * 8;
        printf("ex2.approx: %2d warps: %lld cycles, %.2f ops/clk/SM (%s)\n", warps, h[0], ops / h[0], cudaGetErrorString(e));
    }
    return 0;
}
