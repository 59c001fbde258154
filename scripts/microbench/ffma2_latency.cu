// Micro-benchmark: dependent-chain latency and per-scheduler throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
// One warp per SM sub-partition (128-thread CTA, one CTA per SM); cycles from clock64().  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS, bool PACKED>
__global__ void chain(float* out, long long* cyc, float a, float b, int iters) {
    float2 x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = make_float2(threadIdx.x * 1e-3f + c, c * 0.5f);
    const float2 A = make_float2(a, a * 0.999f), B = make_float2(b, b * 1.001f);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (PACKED) x[c] = __ffma2_rn(x[c], A, B);
                else x[c].x = fmaf(x[c].x, A.x, B.x);
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c].x + x[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS, bool PACKED> void run(const char* name, int threads) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    chain<CHAINS, PACKED><<<148, threads>>>(out, cyc, 0.9999f, 1e-4f, iters);
    chain<CHAINS, PACKED><<<148, threads>>>(out, cyc, 0.9999f, 1e-4f, iters);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 16.0 * CHAINS);
    printf("%-28s threads/CTA %4d chains %d: %.2f cycles per instruction per warp (%.2f per chain step)\n", name, threads, CHAINS, per, per * CHAINS);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, false>("FFMA  dependent", 128); run<1, true>("FFMA2 dependent", 128);
    run<2, false>("FFMA  2 chains", 128);  run<2, true>("FFMA2 2 chains", 128);
    run<4, false>("FFMA  4 chains", 128);  run<4, true>("FFMA2 4 chains", 128);
    run<8, false>("FFMA  8 chains", 128);  run<8, true>("FFMA2 8 chains", 128);
    run<8, false>("FFMA  8 chains", 256);  run<8, true>("FFMA2 8 chains", 256);
    run<8, false>("FFMA  8 chains", 512);  run<8, true>("FFMA2 8 chains", 512);
    return 0;
}
