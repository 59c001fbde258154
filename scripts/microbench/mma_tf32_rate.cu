// Micro-benchmark: issue rate and dependent latency of the warp-level mma.sync.m16n8k8 tf32 (HMMA.1688.F32.TF32) on sm_100a,
// against FFMA on the same schedulers.  One CTA per SM, `threads` threads; CHAINS independent accumulator tiles per warp.
// Question it answers (DESIGN K5): is a 32 x 48 x 192 recurrence step cheaper as 3 x 144 warp MMAs than as 1536 FFMAs per thread?
// Also times the f16 m16n8k16 (HMMA.16816.F32): same 8 cycles per instruction, i.e. twice the flops -- a 2-term fp16 split (3 products per k16) would
// halve the MMA count of the 3xTF32 recurrence at fp16 range limits; not used.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32_rate mma_tf32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_f16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CHAINS, bool F16>
__global__ void mma_chain(float* out, long long* cyc, int iters) {
    float d[CHAINS][4];
    unsigned a[4], b[2];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { d[c][0] = d[c][1] = d[c][2] = d[c][3] = threadIdx.x * 1e-3f + c; }
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + 0.001f * (threadIdx.x + i));
    b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) { if (F16) mma_f16(d[c], a, b); else mma_tf32(d[c], a, b); }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS, bool F16 = false> void run(int threads) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    mma_chain<CHAINS, F16><<<148, threads>>>(out, cyc, iters);
    mma_chain<CHAINS, F16><<<148, threads>>>(out, cyc, iters);
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp = (double)h / (iters * 8.0 * CHAINS);              // cycles per MMA as seen by one warp
    const int warps = threads / 32;
    const double per_sm = per_warp / warps;                                   // cycles per MMA for the whole SM
    const double tflops = 2.0 * 16 * 8 * (F16 ? 16 : 8) / per_sm * 148 * 1.965e9 / 1e12;
    printf("%s threads/CTA %4d chains %2d: %6.2f cycles per MMA per warp, %5.2f per SM  -> %7.1f TFLOP/s (one pass) = %6.1f as 3 products; err %s\n",
           F16 ? "mma.m16n8k16.f16 " : "mma.m16n8k8.tf32 ", threads, CHAINS, per_warp, per_sm, tflops, tflops / 3.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(128); run<2>(128); run<4>(128); run<8>(128);
    run<8>(256); run<8>(384); run<8>(512); run<4>(384); run<2>(384); run<16>(384);
    run<1, true>(128); run<2, true>(128); run<4, true>(128); run<8, true>(128); run<8, true>(384);
    return 0;
}
