// Sequence-persistent BPTT kernels of the PPO learner (ppo2.py:136-197) with the recurrent products on the tensor cores.
//
// Same contract and layouts as lstm_seq_fwd_kernel / lstm_seq_bwd_kernel in policy_kernels.cu (which stay as the FP32-FMA
// variant, IRRL_SEQ_PATH=fma): a CTA owns 32 environments of one tower for ALL T steps, the recurrent state never leaves
// the chip, one launch per layer and direction.  What changes is the per-step product:
//   forward   z[32 x 192]  = hm[32 x 48] . W_h[48 x 192]      (hm = masked h(t-1))
//   backward  dh[32 x 48]  = dz[32 x 192] . W_h^T[192 x 48]
// The FMA version spends 1536 FFMA + 144 LDS.128 per thread and step on it and is bound by the shared-memory pipe (every
// operand of every FMA comes through the LSU).  Here the products are warp-level tensor-core MMAs (mma.sync m16n8k8, tf32
// inputs, fp32 accumulation; HMMA.1688.F32.TF32 in SASS) in the 3xTF32 form of the act kernel (policy_tc_kernels.cu):
// both operands split x = hi + lo, hi the nearest tf32 number, three products lo.hi + hi.lo + hi.hi, error ~2^-22 -- fp32
// results on tf32 hardware.  144 MMAs per warp and step instead of 1536 FFMAs per thread; measured rate of this MMA on
// B200: one per 8 cycles and scheduler (scripts/microbench/mma_tf32_rate.cu), i.e. 1728 SM cycles per CTA step.  The FORWARD product
// goes one step further: a 2-term fp16 split on the f16 m16n8k16 MMA (same issue rate, twice the k: 72 MMAs), see SEQ_FWD_F16 below.
// tcgen05 is the wrong tool for THIS product: the tile per step is 32 x 48 x 192 on a serial dependency (step t+1 needs the
// activations of step t), the weights fit in registers / one shared-memory copy, and the per-step hand-offs of a TMEM
// pipeline (~1.5 k cycles in the act kernel) would cost as much as the product itself.
//
// Fragment ownership (g = lane / 4, q = lane % 4, warp w): the thread owns environments {g, g+8, g+16, g+24} of the tile and
// hidden units {8w+2q, 8w+2q+1} with all four gates -- exactly the C fragments of MMA n-tiles "gate j of units 8w..8w+7"
// -- so the cell update / its derivative run in registers as before.
//   forward:  B fragments (W_h: fp16 leading parts and remainders, 48 registers) live in registers for the whole sequence; A fragments
//             (hm) come from a double-buffered half2 tile that the producers of h write already split (conflict-free fragment
//             reads, one barrier per step).  (-DSEQ_FWD_TF32: tf32 hi parts in registers, bf16 remainders in shared memory, [32][52] tiles.)
//   backward: split-K over warps.  The dz values a thread just computed ARE the A fragments of the k-tile "gate j of this
//             warp's units" (with the k order (2q, 2q+1) <-> (q, q+4), applied to B as well), so dz never goes through shared
//             memory; every warp produces a partial dh[32 x 48] that is summed over the 6 warps through shared memory.
#include <cstdlib>
#include <cuda_fp16.h>
#include "env_device.cuh"
#include "env_kernels.h"

namespace irrl {
namespace seqmma {

constexpr int H = LSTM_H;       // 48
constexpr int G4 = 4 * H;       // 192
constexpr int TM = 32;          // environments per CTA (two m16 tiles)
constexpr int THR = 192;        // 6 warps per environment tile, warp w <-> hidden units 8w .. 8w+7
// Environment tiles per CTA (each tile: 6 warps, its own named barrier and shared-memory block).  Two tiles in one CTA were tried to
// spread the 12 warps evenly over the 4 schedulers (3,3,3,3 instead of 2 x (2,2,1,1)): measured SLOWER at 8192 envs x 750 steps
// (forward 7.97 vs 6.75 ms, backward 8.98 vs 7.94 ms), so one tile per CTA, two CTAs per SM, stays the default.
#ifndef SEQ_GROUPS
#define SEQ_GROUPS 1
#endif
constexpr int GROUPS = SEQ_GROUPS;
__device__ __forceinline__ void tile_sync(int gi) { asm volatile("bar.sync %0, %1;" ::"r"(1 + gi), "n"(THR) : "memory"); }
constexpr int HS_PITCH = 52;    // floats per row of the h tile: (52 g + q) mod 32 distinct over a warp
constexpr int RED_PITCH = 56;   // floats per row of a partial-sum tile: conflict-free float2 stores per half warp

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
// D += A(16x8, row) . B(8x8, col): a0 (g,q) a1 (g+8,q) a2 (g,q+4) a3 (g+8,q+4); b0 (k=q,n=g) b1 (k=q+4,n=g); d0 (g,2q) d1 (g,2q+1) d2 (g+8,2q) d3 (g+8,2q+1)
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split4(const float (&v)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float h = tf32_hi(v[i]); hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(v[i] - h); }
}
#ifndef SEQ_LIBM_ACT     // MUFU forms written out (relative error ~1e-6): sigmoid = rcp(1 + 2^(-x log2 e)), tanh = 1 - 2 rcp(1 + 2^(2 x log2 e)), 4 / 5 instructions each
// (the __expf / __fdividef spellings of the same forms cost about twice the instructions: range fix-ups; libdevice expf / tanhf: 9.5 vs 7.6 ms)
__device__ __forceinline__ float ex2_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sig_(float x) { return rcp_(1.0f + ex2_(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh__(float x) { return fmaf(-2.0f, rcp_(1.0f + ex2_(2.8853900817779268f * x)), 1.0f); }
#else
__device__ __forceinline__ float sig_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float tanh__(float x) { return tanhf(x); }
#endif
// remainder of the tf32 split as bf16 (8 significant bits of a term that is 2^-12 of the value: 2^-20 overall), two per word
__device__ __forceinline__ uint32_t pack_lo(float lo0, float lo1) { return ((__float_as_uint(lo0) + 0x8000u) >> 16) | ((__float_as_uint(lo1) + 0x8000u) & 0xFFFF0000u); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// xw, gates [T,K,N,192] (gate order i,f,o,g); Cs, Hs, HM [T,K,N,48]; keep [T,N]; c0, h0 [K,N,48]; wh [K,48,192]; bias [K,192] or null
#ifndef SEQ_FWD_TF32
#define SEQ_FWD_F16 1          // default; -DSEQ_FWD_TF32 restores the 3xTF32 forward product (6.64 ms against 5.47 ms per launch, error 5.3e-7 against 3.5e-7)
#endif
// ---- SEQ_FWD_F16: the forward product as a 2-term fp16 split (x = x1 + x2, both fp16; x1 x1' + x1 x2' + x2 x1', 2^-22-grade) on the f16
// m16n8k16 MMA, which issues at the same 8 cycles as the tf32 m16n8k8 (scripts/microbench): 72 instead of 144 MMAs per warp and step.
// Only the forward operands (|h| < 1, |W_h| = O(1)) sit safely inside the fp16 range; the remainders of small values land in fp16 subnormals
// (absolute floor 6e-8), which is what bounds the error.
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (a, b) -> half2 of the leading parts and half2 of the remainders
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& x1, uint32_t& x2) {
    const __half2 h1 = __floats2half2_rn(a, b);
    const float2 f1 = __half22float2(h1);
    const __half2 h2 = __floats2half2_rn(a - f1.x, b - f1.y);
    x1 = *reinterpret_cast<const uint32_t*>(&h1); x2 = *reinterpret_cast<const uint32_t*>(&h2);
}
constexpr int HS16_PITCH = 28;   // 32-bit words (half2) per row of the fp16 h tile: (28 g + q) mod 32 distinct over a warp

struct FwdSmem {
    uint32_t Blo[6][4][6][32];        // lo parts of the B fragments [warp][gate j][k-tile][lane] = bf16 (b0, b1)
    float hsh[2][TM][HS_PITCH];       // masked h(t-1), tf32-exact part, row = environment; double buffered over t (one barrier per step)
    float hsl[2][TM][HS_PITCH];       // its remainder   (SEQ_FWD_F16: the same storage viewed as half2 tiles [2][TM][HS16_PITCH])
    float bsm[G4];
};
__global__ void __launch_bounds__(THR * GROUPS, 2 / GROUPS) lstm_seq_fwd_mma_kernel(int T, int K, int N, const float* __restrict__ xw, const float* __restrict__ wh,
                                                                   const float* __restrict__ c0, const float* __restrict__ h0, const float* __restrict__ keep,
                                                                   float* __restrict__ gates, float* __restrict__ Cs, float* __restrict__ Hs,
                                                                   const float* __restrict__ bias, float* __restrict__ HM) {
    extern __shared__ __align__(16) unsigned char seq_smem_f[];
    const int gi = threadIdx.x / THR, tile = blockIdx.x * GROUPS + gi;
    const int tower = blockIdx.y, e0 = tile * TM, t_ = threadIdx.x - gi * THR, w = t_ >> 5, lane = t_ & 31, g = lane >> 2, q = lane & 3;
    if (e0 >= N) return;                                    // the whole group leaves: its barrier is never used
    FwdSmem& S = reinterpret_cast<FwdSmem*>(seq_smem_f)[gi];
    auto& Blo = S.Blo; auto& hsh = S.hsh; auto& hsl = S.hsl; auto& bsm = S.bsm;
    const int u0 = 8 * w + 2 * q;                          // this thread's units u0, u0 + 1
    const float* whk = wh + (size_t)tower * H * G4;
#ifdef SEQ_FWD_F16
    uint32_t b1f[4][3][2], b2f[4][3][2];                   // [gate j][k-tile of 16][b0: k = 2q, 2q+1 | b1: k = 2q+8, 2q+9], leading parts and remainders
    uint32_t (*h1s)[TM][HS16_PITCH] = reinterpret_cast<uint32_t (*)[TM][HS16_PITCH]>(&hsh[0][0][0]);
    uint32_t (*h2s)[TM][HS16_PITCH] = reinterpret_cast<uint32_t (*)[TM][HS16_PITCH]>(&hsl[0][0][0]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int kt = 0; kt < 3; ++kt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int k0 = 16 * kt + 2 * q + 8 * hh, col = j * H + 8 * w + g;
                split_h2(whk[(size_t)k0 * G4 + col], whk[(size_t)(k0 + 1) * G4 + col], b1f[j][kt][hh], b2f[j][kt][hh]);
            }
    (void)Blo;
#else
    uint32_t bh[4][6][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
            const float v0 = whk[(size_t)(8 * kt + q) * G4 + j * H + 8 * w + g], v1 = whk[(size_t)(8 * kt + q + 4) * G4 + j * H + 8 * w + g];
            const float h0_ = tf32_hi(v0), h1_ = tf32_hi(v1);
            bh[j][kt][0] = __float_as_uint(h0_); bh[j][kt][1] = __float_as_uint(h1_);
            Blo[w][j][kt][lane] = pack_lo(v0 - h0_, v1 - h1_);
        }
#endif
    bsm[t_] = bias ? bias[tower * G4 + t_] : 0.f;
    float c[4][2];
    int envs[4];
    auto put_h = [&](int buf, int e, float a, float b) {   // split at the producer: the MMA phase loads ready-made fragments
#ifdef SEQ_FWD_F16
        uint32_t x1, x2; split_h2(a, b, x1, x2);
        h1s[buf][g + 8 * e][u0 >> 1] = x1; h2s[buf][g + 8 * e][u0 >> 1] = x2;
#else
        const float ah = tf32_hi(a), bh_ = tf32_hi(b);
        *reinterpret_cast<float2*>(&hsh[buf][g + 8 * e][u0]) = make_float2(ah, bh_);
        *reinterpret_cast<float2*>(&hsl[buf][g + 8 * e][u0]) = make_float2(a - ah, b - bh_);
#endif
    };
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        envs[e] = min(e0 + g + 8 * e, N - 1);
        const size_t o = ((size_t)tower * N + envs[e]) * H + u0;
        const float k0 = keep[envs[e]];                                     // keep[0][env]
        const float2 cc = *reinterpret_cast<const float2*>(c0 + o), hh = *reinterpret_cast<const float2*>(h0 + o);
        c[e][0] = cc.x * k0; c[e][1] = cc.y * k0;
        put_h(0, e, hh.x * k0, hh.y * k0);
        if (HM) *reinterpret_cast<float2*>(HM + o) = make_float2(hh.x * k0, hh.y * k0);                    // row (t = 0, tower, env)
    }
    float2 xin[4][4]; float kn[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        kn[e] = (T > 1) ? keep[(size_t)N + envs[e]] : 1.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) xin[e][j] = *reinterpret_cast<const float2*>(xw + (((size_t)0 * K + tower) * N + envs[e]) * G4 + j * H + u0);
    }
    tile_sync(gi);
    for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        float acc[2][4][4];                                                 // [m-tile][gate][C fragment]: env e = 2 mt + (c >> 1), unit u0 + (c & 1)
        float knc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            knc[e] = kn[e];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 b2 = *reinterpret_cast<const float2*>(&bsm[j * H + u0]);
                acc[e >> 1][j][2 * (e & 1)] = xin[e][j].x + b2.x; acc[e >> 1][j][2 * (e & 1) + 1] = xin[e][j].y + b2.y;
            }
        }
        if (t + 1 < T) {                                                    // operands of step t+1: in flight behind the whole of step t
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                kn[e] = (t + 2 < T) ? keep[(size_t)(t + 2) * N + envs[e]] : 1.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) xin[e][j] = *reinterpret_cast<const float2*>(xw + (((size_t)(t + 1) * K + tower) * N + envs[e]) * G4 + j * H + u0);
            }
        }
#ifdef SEQ_FWD_F16
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            uint32_t a1[2][4], a2[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {                            // (g, 2q..) (g+8, 2q..) (g, 2q+8..) (g+8, 2q+8..), two consecutive k per word
                a1[mt][0] = h1s[buf][16 * mt + g][8 * kt + q];     a1[mt][1] = h1s[buf][16 * mt + g + 8][8 * kt + q];
                a1[mt][2] = h1s[buf][16 * mt + g][8 * kt + q + 4]; a1[mt][3] = h1s[buf][16 * mt + g + 8][8 * kt + q + 4];
                a2[mt][0] = h2s[buf][16 * mt + g][8 * kt + q];     a2[mt][1] = h2s[buf][16 * mt + g + 8][8 * kt + q];
                a2[mt][2] = h2s[buf][16 * mt + g][8 * kt + q + 4]; a2[mt][3] = h2s[buf][16 * mt + g + 8][8 * kt + q + 4];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_f16(acc[mt][j], a2[mt], b1f[j][kt][0], b1f[j][kt][1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_f16(acc[mt][j], a1[mt], b2f[j][kt][0], b2f[j][kt][1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_f16(acc[mt][j], a1[mt], b1f[j][kt][0], b1f[j][kt][1]);
        }
#else
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                ah[mt][0] = __float_as_uint(hsh[buf][16 * mt + g][8 * kt + q]);     ah[mt][1] = __float_as_uint(hsh[buf][16 * mt + g + 8][8 * kt + q]);
                ah[mt][2] = __float_as_uint(hsh[buf][16 * mt + g][8 * kt + q + 4]); ah[mt][3] = __float_as_uint(hsh[buf][16 * mt + g + 8][8 * kt + q + 4]);
                al[mt][0] = __float_as_uint(hsl[buf][16 * mt + g][8 * kt + q]);     al[mt][1] = __float_as_uint(hsl[buf][16 * mt + g + 8][8 * kt + q]);
                al[mt][2] = __float_as_uint(hsl[buf][16 * mt + g][8 * kt + q + 4]); al[mt][3] = __float_as_uint(hsl[buf][16 * mt + g + 8][8 * kt + q + 4]);
            }
            uint32_t bl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bl[j] = Blo[w][j][kt][lane];
            // three passes over 8 independent accumulators: consecutive MMAs never depend on each other
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma(acc[mt][j], al[mt], bh[j][kt][0], bh[j][kt][1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma(acc[mt][j], ah[mt], bl[j] << 16, bl[j] & 0xFFFF0000u);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma(acc[mt][j], ah[mt], bh[j][kt][0], bh[j][kt][1]);
        }
#endif
        // per m-tile: its 4 cells as one block of independent arithmetic, then the stores (all 8 at once costs spills at 168 registers).
        // Rows past N alias row N-1 (envs[] is clamped): those threads compute bit-identical values from identical inputs, so the stores
        // need no predicate.
        const size_t rowbase = ((size_t)t * K + tower) * N;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            float gI[2][2], gF[2][2], gO[2][2], gG[2][2], hN[2][2];
#pragma unroll
            for (int eh = 0; eh < 2; ++eh)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int e = 2 * mt + eh, ci = 2 * eh + u;
                    gI[eh][u] = sig_(acc[mt][0][ci]); gF[eh][u] = sig_(acc[mt][1][ci]); gO[eh][u] = sig_(acc[mt][2][ci]); gG[eh][u] = tanh__(acc[mt][3][ci]);
                    c[e][u] = gF[eh][u] * c[e][u] + gI[eh][u] * gG[eh][u];
                    hN[eh][u] = gO[eh][u] * tanh__(c[e][u]);
                }
#pragma unroll
            for (int eh = 0; eh < 2; ++eh) {
                const int e = 2 * mt + eh;
                const size_t row = rowbase + envs[e];
                float* gr = gates + row * G4 + u0;
                *reinterpret_cast<float2*>(gr) = make_float2(gI[eh][0], gI[eh][1]); *reinterpret_cast<float2*>(gr + H) = make_float2(gF[eh][0], gF[eh][1]);
                *reinterpret_cast<float2*>(gr + 2 * H) = make_float2(gO[eh][0], gO[eh][1]); *reinterpret_cast<float2*>(gr + 3 * H) = make_float2(gG[eh][0], gG[eh][1]);
                *reinterpret_cast<float2*>(Cs + row * H + u0) = make_float2(c[e][0], c[e][1]);
                *reinterpret_cast<float2*>(Hs + row * H + u0) = make_float2(hN[eh][0], hN[eh][1]);
                c[e][0] *= knc[e]; c[e][1] *= knc[e];                       // SB lstm(): c *= 1-m ; h *= 1-m before the next cell
                put_h(buf ^ 1, e, hN[eh][0] * knc[e], hN[eh][1] * knc[e]);  // the other buffer: nobody reads it before the barrier below
                if (HM && t + 1 < T) *reinterpret_cast<float2*>(HM + (rowbase + (size_t)K * N + envs[e]) * H + u0) = make_float2(hN[eh][0] * knc[e], hN[eh][1] * knc[e]);
            }
        }
        // (An in-warp software pipeline -- MMAs of m-tile 1 interleaved with the cell arithmetic of m-tile 0 -- was measured: 6.83 vs 6.64 ms.)
        tile_sync(gi);                                                    // h(t) complete; everybody has also finished reading h(t-1)
    }
}

// dH [T,K,N,48], gates / dz [T,K,N,192], Cs [T,K,N,48], keep [T,N], c0 [K,N,48], wh [K,48,192]; db_part [CTAs,K,192] or null
constexpr int BWD_SMEM = 6 * 4 * 6 * 32 * (8 + 4) + 6 * TM * RED_PITCH * 4;     // B hi (float2) + B lo (2 x bf16) + partial sums = 98 304 B
__global__ void __launch_bounds__(THR * GROUPS, 2 / GROUPS) lstm_seq_bwd_mma_kernel(int T, int K, int N, const float* __restrict__ dH, const float* __restrict__ wh,
                                                                   const float* __restrict__ c0, const float* __restrict__ keep, const float* __restrict__ gates,
                                                                   const float* __restrict__ Cs, float* __restrict__ dz, float* __restrict__ db_part) {
    extern __shared__ __align__(16) unsigned char seq_smem_b[];
    const int gi = threadIdx.x / THR, tile = blockIdx.x * GROUPS + gi;
    unsigned char* smem = seq_smem_b + (size_t)gi * BWD_SMEM;
#ifdef SEQ_BWD_F16
    uint2 (*B1s)[2][6][32] = reinterpret_cast<uint2 (*)[2][6][32]>(smem);                                    // [warp][k-tile = gate pair][n-tile][lane] = (b0, b1) half2 leading parts
    uint2 (*B2s)[2][6][32] = reinterpret_cast<uint2 (*)[2][6][32]>(smem + 6 * 2 * 6 * 32 * 8);                // their remainders
#endif
    float2 (*Bhi)[4][6][32] = reinterpret_cast<float2 (*)[4][6][32]>(smem);                                  // [warp][k-tile = gate j][n-tile][lane] = (b0, b1), tf32-exact
    uint32_t (*Blo)[4][6][32] = reinterpret_cast<uint32_t (*)[4][6][32]>(smem + 6 * 4 * 6 * 32 * 8);         // remainders as two bf16 (b0 low half, b1 high half)
    float (*red)[TM][RED_PITCH] = reinterpret_cast<float (*)[TM][RED_PITCH]>(smem + 6 * 4 * 6 * 32 * 12);    // [warp] partial dh[32][48]
    const int tower = blockIdx.y, e0 = tile * TM, t_ = threadIdx.x - gi * THR, w = t_ >> 5, lane = t_ & 31, g = lane >> 2, q = lane & 3;
    if (e0 >= N) return;
    const int u0 = 8 * w + 2 * q;
    const float* whk = wh + (size_t)tower * H * G4;
    // B = W_h^T restricted to this warp's 32 dz columns: k-tile j = gate j, MMA k index q <-> column j*48 + u0, q+4 <-> column j*48 + u0 + 1; n = 8 nt + g
#ifdef SEQ_BWD_F16
    // f16 m16n8k16: a k-tile is a gate PAIR (2 kp, 2 kp + 1): MMA k = 2q, 2q+1 <-> gate 2 kp, units u0, u0+1; k = 2q+8, 2q+9 <-> gate 2 kp + 1
#pragma unroll
    for (int kp = 0; kp < 2; ++kp)
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            const float2 v0 = *reinterpret_cast<const float2*>(whk + (size_t)(8 * nt + g) * G4 + (2 * kp) * H + u0);
            const float2 v1 = *reinterpret_cast<const float2*>(whk + (size_t)(8 * nt + g) * G4 + (2 * kp + 1) * H + u0);
            uint2 x1, x2; split_h2(v0.x, v0.y, x1.x, x2.x); split_h2(v1.x, v1.y, x1.y, x2.y);
            B1s[w][kp][nt][lane] = x1; B2s[w][kp][nt][lane] = x2;
        }
    (void)Bhi; (void)Blo;
#else
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            const float2 v = *reinterpret_cast<const float2*>(whk + (size_t)(8 * nt + g) * G4 + j * H + u0);
            const float hx = tf32_hi(v.x), hy = tf32_hi(v.y);
            Bhi[w][j][nt][lane] = make_float2(hx, hy);
            const uint32_t lx = (__float_as_uint(v.x - hx) + 0x8000u) >> 16, ly = (__float_as_uint(v.y - hy) + 0x8000u) & 0xFFFF0000u;
            Blo[w][j][nt][lane] = lx | ly;
        }
#endif
    int envs[4]; bool valid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { envs[e] = min(e0 + g + 8 * e, N - 1); valid[e] = (e0 + g + 8 * e) < N; }
    float dbacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float ch[4][2], cc[4][2];
#pragma unroll
    for (int e = 0; e < 4; ++e) { ch[e][0] = ch[e][1] = cc[e][0] = cc[e][1] = 0.f; }
    // Per-step operands are loaded at the top of their step: a one-step-ahead register prefetch (64 registers in flight through the MMA
    // phase) forced spills in the hot loop (ncu: STL 22 % of the stall samples).  Their lines are pulled into L2 one step ahead instead
    // (prefetch.global.L2 behind the element-wise part), so the loads hit L2; c(t-1) and keep(t) are carried over from the step before.
    struct In { float2 dh, cp, gi, gf, go, gg; float kt; };
    float2 ccur[4]; float kuc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { ccur[e] = *reinterpret_cast<const float2*>(Cs + (((size_t)(T - 1) * K + tower) * N + envs[e]) * H + u0); kuc[e] = 0.f; }
    auto load_step = [&](int t, In* in) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const size_t row = ((size_t)t * K + tower) * N + envs[e];
            in[e].kt = keep[(size_t)t * N + envs[e]];
            in[e].dh = *reinterpret_cast<const float2*>(dH + row * H + u0);
            in[e].cp = (t > 0) ? *reinterpret_cast<const float2*>(Cs + (((size_t)(t - 1) * K + tower) * N + envs[e]) * H + u0)
                               : *reinterpret_cast<const float2*>(c0 + ((size_t)tower * N + envs[e]) * H + u0);
            const float* gr = gates + row * G4 + u0;
            in[e].gi = *reinterpret_cast<const float2*>(gr); in[e].gf = *reinterpret_cast<const float2*>(gr + H);
            in[e].go = *reinterpret_cast<const float2*>(gr + 2 * H); in[e].gg = *reinterpret_cast<const float2*>(gr + 3 * H);
        }
    };
    auto prefetch_step = [&](int t) {     // one request per 32-byte sector: the four lanes q = 0..3 share it
        if (q == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const size_t row = ((size_t)t * K + tower) * N + envs[e];
                prefetch_l2(dH + row * H + u0);
                if (t > 0) prefetch_l2(Cs + (((size_t)(t - 1) * K + tower) * N + envs[e]) * H + u0);
                const float* gr = gates + row * G4 + u0;
                prefetch_l2(gr); prefetch_l2(gr + H); prefetch_l2(gr + 2 * H); prefetch_l2(gr + 3 * H);
            }
        }
    };
    tile_sync(gi);
#ifdef SEQ_BWD_LATE_LOADS
    In cur[4];
    load_step(T - 1, cur);
#endif
    for (int t = T - 1; t >= 0; --t) {
#ifndef SEQ_BWD_LATE_LOADS
        In cur[4];
        load_step(t, cur);
#endif
        float dzv[4][2][4];                                                 // [env e][unit u][gate]
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const size_t row = ((size_t)t * K + tower) * N + envs[e];
            const float ku = kuc[e], kt = cur[e].kt;
            const float dhv[2] = {cur[e].dh.x + ch[e][0] * ku, cur[e].dh.y + ch[e][1] * ku}, cv[2] = {ccur[e].x, ccur[e].y}, cpm[2] = {cur[e].cp.x * kt, cur[e].cp.y * kt};
            const float iv[2] = {cur[e].gi.x, cur[e].gi.y}, fv[2] = {cur[e].gf.x, cur[e].gf.y}, ov[2] = {cur[e].go.x, cur[e].go.y}, gv[2] = {cur[e].gg.x, cur[e].gg.y};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float tc = tanh__(cv[u]);
                const float dc = cc[e][u] * ku + dhv[u] * ov[u] * (1.f - tc * tc);
                dzv[e][u][0] = dc * gv[u] * iv[u] * (1.f - iv[u]); dzv[e][u][1] = dc * cpm[u] * fv[u] * (1.f - fv[u]);
                dzv[e][u][2] = dhv[u] * tc * ov[u] * (1.f - ov[u]); dzv[e][u][3] = dc * iv[u] * (1.f - gv[u] * gv[u]);
                cc[e][u] = dc * fv[u];
            }
            {   // rows past N alias row N-1 with bit-identical values (envs[] is clamped): no predicate on the stores, a 0 / 1 weight on the bias sums
                float* dr = dz + row * G4 + u0;
                const float vw = valid[e] ? 1.f : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) { *reinterpret_cast<float2*>(dr + j * H) = make_float2(dzv[e][0][j], dzv[e][1][j]); dbacc[0][j] = fmaf(vw, dzv[e][0][j], dbacc[0][j]); dbacc[1][j] = fmaf(vw, dzv[e][1][j], dbacc[1][j]); }
            }
            ccur[e] = cur[e].cp; kuc[e] = kt;                               // c(t-1) and keep(t) are what step t-1 calls c and keep(t+1)
        }
        if (t > 0) prefetch_step(t - 1);
#ifdef SEQ_BWD_F16
        // 2-term fp16 split on the f16 m16n8k16 MMA (72 instead of 144 MMAs).  Gradients have no fixed range, so the warp scales its 32 x 32 dz block
        // by a power of two that puts the largest magnitude at 2^13..2^14 (exact), and un-scales its partial sums (exact): elements far below the
        // block maximum keep an ABSOLUTE accuracy of ~2^-39 of that maximum, the rest 2^-22 relative.
        float sf, isf;
        {
            float m = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < 4; ++j) m = fmaxf(m, fabsf(dzv[e][u][j]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const int ex = (int)((__float_as_uint(m) >> 23) & 0xffu) - 127;             // floor(log2 m); m = 0 or denormal -> -127
            const int sh = max(-100, min(100, 13 - ex));
            sf = __uint_as_float((uint32_t)(sh + 127) << 23); isf = __uint_as_float((uint32_t)(127 - sh) << 23);
        }
        uint32_t a1[2][2][4], a2[2][2][4];                                  // [gate pair][m-tile][fragment]: (g, gate 2kp) (g+8, gate 2kp) (g, gate 2kp+1) (g+8, gate 2kp+1), units (u0, u0+1) per word
#pragma unroll
        for (int kp = 0; kp < 2; ++kp)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                split_h2(dzv[2 * mt][0][2 * kp] * sf, dzv[2 * mt][1][2 * kp] * sf, a1[kp][mt][0], a2[kp][mt][0]);
                split_h2(dzv[2 * mt + 1][0][2 * kp] * sf, dzv[2 * mt + 1][1][2 * kp] * sf, a1[kp][mt][1], a2[kp][mt][1]);
                split_h2(dzv[2 * mt][0][2 * kp + 1] * sf, dzv[2 * mt][1][2 * kp + 1] * sf, a1[kp][mt][2], a2[kp][mt][2]);
                split_h2(dzv[2 * mt + 1][0][2 * kp + 1] * sf, dzv[2 * mt + 1][1][2 * kp + 1] * sf, a1[kp][mt][3], a2[kp][mt][3]);
            }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float part[2][3][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3) part[mt][n3][0] = part[mt][n3][1] = part[mt][n3][2] = part[mt][n3][3] = 0.f;
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {
                uint2 b1[3], b2[3];
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3) { b1[n3] = B1s[w][kp][3 * p + n3][lane]; b2[n3] = B2s[w][kp][3 * p + n3][lane]; }
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma_f16(part[mt][n3], a2[kp][mt], b1[n3].x, b1[n3].y);
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma_f16(part[mt][n3], a1[kp][mt], b2[n3].x, b2[n3].y);
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma_f16(part[mt][n3], a1[kp][mt], b1[n3].x, b1[n3].y);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3) {
                    const int col = 8 * (3 * p + n3) + 2 * q;
                    *reinterpret_cast<float2*>(&red[w][16 * mt + g][col]) = make_float2(part[mt][n3][0] * isf, part[mt][n3][1] * isf);
                    *reinterpret_cast<float2*>(&red[w][16 * mt + g + 8][col]) = make_float2(part[mt][n3][2] * isf, part[mt][n3][3] * isf);
                }
        }
#else
        // partial dh of this warp: A fragment of (m-tile mt, k-tile j) = dz values this thread already holds
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float part[2][3][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3) part[mt][n3][0] = part[mt][n3][1] = part[mt][n3][2] = part[mt][n3][3] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t ah[2][4], al[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float v[4] = {dzv[2 * mt][0][j], dzv[2 * mt + 1][0][j], dzv[2 * mt][1][j], dzv[2 * mt + 1][1][j]};      // (g,q) (g+8,q) (g,q+4) (g+8,q+4)
                    split4(v, ah[mt], al[mt]);
                }
                float2 bh[3]; uint32_t bl[3];
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3) { bh[n3] = Bhi[w][j][3 * p + n3][lane]; bl[n3] = Blo[w][j][3 * p + n3][lane]; }
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma(part[mt][n3], al[mt], __float_as_uint(bh[n3].x), __float_as_uint(bh[n3].y));
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma(part[mt][n3], ah[mt], bl[n3] << 16, bl[n3] & 0xFFFF0000u);
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma(part[mt][n3], ah[mt], __float_as_uint(bh[n3].x), __float_as_uint(bh[n3].y));
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n3 = 0; n3 < 3; ++n3) {
                    const int col = 8 * (3 * p + n3) + 2 * q;
                    *reinterpret_cast<float2*>(&red[w][16 * mt + g][col]) = make_float2(part[mt][n3][0], part[mt][n3][1]);
                    *reinterpret_cast<float2*>(&red[w][16 * mt + g + 8][col]) = make_float2(part[mt][n3][2], part[mt][n3][3]);
                }
        }
#endif
        tile_sync(gi);
#ifdef SEQ_BWD_LATE_LOADS                                                   // variant: register loads behind the reduction (spills 192 B at 168 registers)
        if (t > 0) load_step(t - 1, cur);
#endif
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 s = *reinterpret_cast<const float2*>(&red[0][g + 8 * e][u0]);
#pragma unroll
            for (int w2 = 1; w2 < 6; ++w2) { const float2 v = *reinterpret_cast<const float2*>(&red[w2][g + 8 * e][u0]); s.x += v.x; s.y += v.y; }
            ch[e][0] = s.x; ch[e][1] = s.y;
        }
        tile_sync(gi);
    }
    if (db_part) {   // the 8 environment groups of a unit pair are lanes q, q+4, .., q+28: fixed-order butterfly, one partial per CTA (summed afterwards: deterministic)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = dbacc[u][j];
                v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (g == 0) db_part[((size_t)tile * K + tower) * G4 + j * H + u0 + u] = v;
            }
    }
}

}  // namespace seqmma

// 0 (default) = tensor-core kernels of this file, 1 = the FP32-FMA kernels of policy_kernels.cu (A/B and regression reference).
// Initial value from IRRL_SEQ_PATH=fma; irrl_lstm_seq_set_path switches in-process.
static int g_seq_path = [] { const char* e = getenv("IRRL_SEQ_PATH"); return (e && e[0] == 'f') ? 1 : 0; }();
int lstm_seq_set_path(int path) { const int prev = g_seq_path; if (path == 0 || path == 1) g_seq_path = path; return prev; }
static bool seq_use_fma() { return g_seq_path == 1; }

void launch_lstm_seq_fwd(int T, int K, int N, const float* xw, const float* wh, const float* c0, const float* h0, const float* keep, float* gates, float* Cs,
                         float* Hs, const float* bias, float* HM, cudaStream_t st) {
    if (seq_use_fma()) { launch_lstm_seq_fwd_fma(T, K, N, xw, wh, c0, h0, keep, gates, Cs, Hs, bias, HM, st); return; }
    const int tiles = (N + seqmma::TM - 1) / seqmma::TM;
    dim3 grid((tiles + seqmma::GROUPS - 1) / seqmma::GROUPS, K);
    constexpr int smem = sizeof(seqmma::FwdSmem) * seqmma::GROUPS;
    {   // the attribute is per device (a process may hold envs / policies on several GPUs through the C ABI)
        static unsigned long long configured_devices = 0ull; int dev = 0; cudaGetDevice(&dev);
        if (dev >= 64 || !((configured_devices >> dev) & 1ull)) {
            cudaFuncSetAttribute(seqmma::lstm_seq_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (dev < 64) configured_devices |= 1ull << dev;
        }
    }
    seqmma::lstm_seq_fwd_mma_kernel<<<grid, seqmma::THR * seqmma::GROUPS, smem, st>>>(T, K, N, xw, wh, c0, h0, keep, gates, Cs, Hs, bias, HM);
}
void launch_lstm_seq_bwd(int T, int K, int N, const float* dH, const float* wh, const float* c0, const float* keep, const float* gates, const float* Cs,
                         float* dz, float* db_part, cudaStream_t st) {
    if (seq_use_fma()) { launch_lstm_seq_bwd_fma(T, K, N, dH, wh, c0, keep, gates, Cs, dz, db_part, st); return; }
    const int tiles = (N + seqmma::TM - 1) / seqmma::TM;
    dim3 grid((tiles + seqmma::GROUPS - 1) / seqmma::GROUPS, K);
    {   // the attribute is per device (a process may hold envs / policies on several GPUs through the C ABI)
        static unsigned long long configured_devices = 0ull; int dev = 0; cudaGetDevice(&dev);
        if (dev >= 64 || !((configured_devices >> dev) & 1ull)) {
            cudaFuncSetAttribute(seqmma::lstm_seq_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, seqmma::BWD_SMEM * seqmma::GROUPS);
            if (dev < 64) configured_devices |= 1ull << dev;
        }
    }
    seqmma::lstm_seq_bwd_mma_kernel<<<grid, seqmma::THR * seqmma::GROUPS, seqmma::BWD_SMEM * seqmma::GROUPS, st>>>(T, K, N, dH, wh, c0, keep, gates, Cs, dz, db_part);
}

}  // namespace irrl
