// The non-recurrent products of the PPO learner's BPTT (ppo2.py:136-197 over run_bp_v5.py:143-166) on the tensor cores:
//   proj_rows   Y[t,k,n,:] = X[t,(k),n,:] . B_k          input projections x W_x of all T steps (Kin 35 / 48 -> 192 gate columns)
//                                                          and their input gradient dz W_x^T (192 -> 48), B_k = W_k or W_k^T
//   gram_rows   G_k = sum_{t,n} X[t,(k),n,:]^T D[t,k,n,:]   weight gradients dW_x, dW_h (Kin x 192, reduction over T*N ~ 6 M rows)
// Every one of them is a pass over a multi-GB time-major tensor ([T,K,N,192] is 9.4 GB at 8192 envs x 750 steps) with a tiny
// second operand: the bound is HBM, provided the arithmetic keeps up.  cuBLAS runs them as FP32 SIMT GEMMs at 5 - 8 ms each
// (HBM time 1.5 - 2 ms).  Here: warp-level tensor-core MMAs (mma.sync m16n8k8 tf32, fp32 accumulate) in the 3xTF32 form of the
// act / sequence kernels (operands split hi + lo on the fly, three products; fp32-grade results), tiles streamed with cp.async
// through double-buffered shared memory, the small operand packed once per CTA in fragment order.
// Rows are addressed as (t, tower k, env n); X may lack the tower dimension (the observation is shared by both towers).
#include <algorithm>
#include <cstdlib>
#include "env_device.cuh"
#include "env_kernels.h"

namespace irrl {
namespace lgemm {

constexpr int THR = 256;
constexpr int NSTAGE = 3;       // cp.async stages: tiles i+1 and i+2 are in flight while tile i is multiplied; ONE barrier per tile

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split(float v, uint32_t& hi, uint32_t& lo) { const float h = tf32_hi(v); hi = __float_as_uint(h); lo = __float_as_uint(v - h); }
__device__ __forceinline__ uint32_t pack_lo(float lo0, float lo1) { return ((__float_as_uint(lo0) + 0x8000u) >> 16) | ((__float_as_uint(lo1) + 0x8000u) & 0xFFFF0000u); }
__device__ __forceinline__ void cp16(void* s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g)); }
__device__ __forceinline__ void cp4(void* s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g)); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct RowArgs {
    const float* X; int x_cols; long long x_t_stride, x_k_stride;     // row (t,k,n) of X starts at X + (t*x_t_stride + k*x_k_stride + n) * x_cols
    const float* W; int w_ld, w_trans; long long w_k_stride;          // B_k[kk][nn] = w_trans ? W_k[nn*w_ld + kk] : W_k[kk*w_ld + nn]
    float* Y;                                                         // proj: [T,K,N,NT*8]; gram: partial sums [gridDim.x][K][MT*16][192]
    const float* D;                                                   // gram: [T,K,N,192]
    int T, K, N;
};

// copies rows [n0, n0 + ROWS) of (t, k) into a shared tile of row pitch PITCH floats; rows past N are zero-filled when ZERO is set
template <int ROWS, int PITCH, bool ZERO>
__device__ __forceinline__ void load_rows(float* tile, const float* src_base, int cols, int n0, int N, bool vec) {
    const int t_ = threadIdx.x;
    if (vec) {
        const int cpr = cols >> 2;
        for (int i = t_; i < ROWS * cpr; i += THR) {
            const int r = i / cpr, c = i - r * cpr;
            if (n0 + r < N) cp16(tile + r * PITCH + 4 * c, src_base + (size_t)r * cols + 4 * c);
            else if (ZERO) *reinterpret_cast<float4*>(tile + r * PITCH + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int i = t_; i < ROWS * cols; i += THR) {
            const int r = i / cols, c = i - r * cols;
            if (n0 + r < N) cp4(tile + r * PITCH + c, src_base + (size_t)r * cols + c);
            else if (ZERO) tile[r * PITCH + c] = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- proj_rows
// CTA tile = 64 rows x (NT*8) columns; warp w: m-tile w & 3 (16 rows), column half w >> 2 (NT/2 n-tiles).  KT k-tiles (Kin padded to 8).
template <int KT, int NT>
struct ProjSmem {
    static constexpr int PITCH = KT * 8 + 4;                           // (4 g + q) mod 32 distinct: conflict-free A fragments
    float2 Bhi[KT][NT][32]; uint32_t Blo[KT][NT][32];
    float Xs[NSTAGE][64][PITCH];
};
template <int KT, int NT>
__global__ void __launch_bounds__(THR, (KT <= 6) ? 2 : 1) proj_rows_kernel(const __grid_constant__ RowArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using S = ProjSmem<KT, NT>;
    S& s = *reinterpret_cast<S*>(smem_raw);
    constexpr int PITCH = S::PITCH, NH = NT / 2, GRP = (NH % 4 == 0) ? 4 : 3;
    const int t_ = threadIdx.x, w = t_ >> 5, lane = t_ & 31, g = lane >> 2, q = lane & 3, mw = w & 3, ch = w >> 2;
    const int k = blockIdx.y;
    const float* Wk = A.W + (size_t)k * A.w_k_stride;
    for (int i = t_; i < KT * NT * 32; i += THR) {
        const int ln = i & 31, nt = (i >> 5) % NT, kt = (i >> 5) / NT, kk0 = 8 * kt + (ln & 3), kk1 = kk0 + 4, nn = 8 * nt + (ln >> 2);
        const float v0 = kk0 < A.x_cols ? (A.w_trans ? Wk[(size_t)nn * A.w_ld + kk0] : Wk[(size_t)kk0 * A.w_ld + nn]) : 0.f;
        const float v1 = kk1 < A.x_cols ? (A.w_trans ? Wk[(size_t)nn * A.w_ld + kk1] : Wk[(size_t)kk1 * A.w_ld + nn]) : 0.f;
        const float h0 = tf32_hi(v0), h1 = tf32_hi(v1);
        s.Bhi[kt][nt][ln] = make_float2(h0, h1); s.Blo[kt][nt][ln] = pack_lo(v0 - h0, v1 - h1);
    }
    for (int i = t_; i < NSTAGE * 64 * PITCH; i += THR) (&s.Xs[0][0][0])[i] = 0.f;      // pad columns stay zero (the copies never touch them)
    __syncthreads();
    const int tiles_per_t = (A.N + 63) / 64, tiles = A.T * tiles_per_t;
    const bool vec = (A.x_cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(A.X) & 15) == 0);
    auto issue = [&](int tile, int stage) {
        const int t = tile / tiles_per_t, n0 = (tile - t * tiles_per_t) * 64;
        load_rows<64, PITCH, false>(&s.Xs[stage][0][0], A.X + ((size_t)t * A.x_t_stride + (size_t)k * A.x_k_stride + n0) * A.x_cols, A.x_cols, n0, A.N, vec);
        cp_commit();
    };
    int tile = blockIdx.x, stage = 0;
    if (tile < tiles) issue(tile, 0); else cp_commit();
    if (tile + (int)gridDim.x < tiles) issue(tile + gridDim.x, 1); else cp_commit();
    for (; tile < tiles; tile += gridDim.x, stage = (stage + 1) % NSTAGE) {
        cp_wait<1>();                                                       // every group but the newest has landed: this tile is in shared memory
        __syncthreads();                                                    // ... for every thread, and everybody has finished the previous tile
        const int ahead = tile + 2 * (int)gridDim.x;                        // refill the stage the previous tile was multiplied from
        if (ahead < tiles) issue(ahead, (stage + 2) % NSTAGE); else cp_commit();
        float acc[NH][4];
#pragma unroll
        for (int i = 0; i < NH; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        const float (*X)[PITCH] = s.Xs[stage];
#pragma unroll 2
        for (int kt = 0; kt < KT; ++kt) {
            uint32_t ah[4], al[4];
            split(X[16 * mw + g][8 * kt + q], ah[0], al[0]);     split(X[16 * mw + g + 8][8 * kt + q], ah[1], al[1]);
            split(X[16 * mw + g][8 * kt + q + 4], ah[2], al[2]); split(X[16 * mw + g + 8][8 * kt + q + 4], ah[3], al[3]);
#pragma unroll
            for (int n0 = 0; n0 < NH; n0 += GRP) {
                float2 bh[GRP]; uint32_t bl[GRP];
#pragma unroll
                for (int i = 0; i < GRP; ++i) { bh[i] = s.Bhi[kt][ch * NH + n0 + i][lane]; bl[i] = s.Blo[kt][ch * NH + n0 + i][lane]; }
#pragma unroll
                for (int i = 0; i < GRP; ++i) mma(acc[n0 + i], al, __float_as_uint(bh[i].x), __float_as_uint(bh[i].y));
#pragma unroll
                for (int i = 0; i < GRP; ++i) mma(acc[n0 + i], ah, bl[i] << 16, bl[i] & 0xFFFF0000u);
#pragma unroll
                for (int i = 0; i < GRP; ++i) mma(acc[n0 + i], ah, __float_as_uint(bh[i].x), __float_as_uint(bh[i].y));
            }
        }
        const int t = tile / tiles_per_t, n0 = (tile - t * tiles_per_t) * 64;
        float* yb = A.Y + (((size_t)t * A.K + k) * A.N) * (NT * 8);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int n = n0 + 16 * mw + g + 8 * half;
            if (n < A.N) {
#pragma unroll
                for (int i = 0; i < NH; ++i) *reinterpret_cast<float2*>(yb + (size_t)n * (NT * 8) + 8 * (ch * NH + i) + 2 * q) = make_float2(acc[i][2 * half], acc[i][2 * half + 1]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- gram_rows
// G[m][n] = sum_rows X[row][m] D[row][n]: MMA M = feature m (MT m-tiles), N = 192 gate columns (warp w: n-tiles 3w .. 3w+2), K = rows.
// CTA tile = 32 rows per stage; every CTA keeps its accumulators over all its tiles and writes one partial sum (summed afterwards in a
// fixed order: deterministic).
template <int MT>
struct GramSmem {
    static constexpr int XP = MT * 16 + 8, DP = 200;                   // (24 q + g) / (8 q + g) mod 32 distinct: conflict-free fragments
    float Xs[NSTAGE][32][XP]; float Ds[NSTAGE][32][DP];
};
template <int MT>
__global__ void __launch_bounds__(THR, 2) gram_rows_kernel(const __grid_constant__ RowArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using S = GramSmem<MT>;
    S& s = *reinterpret_cast<S*>(smem_raw);
    constexpr int XP = S::XP, DP = S::DP;
    const int t_ = threadIdx.x, w = t_ >> 5, lane = t_ & 31, g = lane >> 2, q = lane & 3;
    const int k = blockIdx.y;
    for (int i = t_; i < NSTAGE * 32 * XP; i += THR) (&s.Xs[0][0][0])[i] = 0.f;          // pad columns (features past x_cols) stay zero
    __syncthreads();
    const int tiles_per_t = (A.N + 31) / 32, tiles = A.T * tiles_per_t;
    const bool xvec = (A.x_cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(A.X) & 15) == 0);
    auto issue = [&](int tile, int stage) {
        const int t = tile / tiles_per_t, n0 = (tile - t * tiles_per_t) * 32;
        load_rows<32, XP, true>(&s.Xs[stage][0][0], A.X + ((size_t)t * A.x_t_stride + (size_t)k * A.x_k_stride + n0) * A.x_cols, A.x_cols, n0, A.N, xvec);
        load_rows<32, DP, true>(&s.Ds[stage][0][0], A.D + (((size_t)t * A.K + k) * A.N + n0) * 192, 192, n0, A.N, true);
        cp_commit();
    };
    float acc[MT][3][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[m][i][0] = acc[m][i][1] = acc[m][i][2] = acc[m][i][3] = 0.f;
    int tile = blockIdx.x, stage = 0;
    if (tile < tiles) issue(tile, 0); else cp_commit();
    if (tile + (int)gridDim.x < tiles) issue(tile + gridDim.x, 1); else cp_commit();
    for (; tile < tiles; tile += gridDim.x, stage = (stage + 1) % NSTAGE) {
        cp_wait<1>();                                                       // every group but the newest has landed: this tile is in shared memory
        __syncthreads();                                                    // ... for every thread, and everybody has finished the previous tile
        const int ahead = tile + 2 * (int)gridDim.x;                        // refill the stage the previous tile was multiplied from
        if (ahead < tiles) issue(ahead, (stage + 2) % NSTAGE); else cp_commit();
        const float (*X)[XP] = s.Xs[stage]; const float (*D)[DP] = s.Ds[stage];
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            uint32_t ah[MT][4], al[MT][4], bh[3][2], bl[3][2];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                split(X[8 * kt + q][16 * m + g], ah[m][0], al[m][0]);     split(X[8 * kt + q][16 * m + g + 8], ah[m][1], al[m][1]);
                split(X[8 * kt + q + 4][16 * m + g], ah[m][2], al[m][2]); split(X[8 * kt + q + 4][16 * m + g + 8], ah[m][3], al[m][3]);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) { split(D[8 * kt + q][8 * (3 * w + i) + g], bh[i][0], bl[i][0]); split(D[8 * kt + q + 4][8 * (3 * w + i) + g], bh[i][1], bl[i][1]); }
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int i = 0; i < 3; ++i) mma(acc[m][i], al[m], bh[i][0], bh[i][1]);
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int i = 0; i < 3; ++i) mma(acc[m][i], ah[m], bl[i][0], bl[i][1]);
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int i = 0; i < 3; ++i) mma(acc[m][i], ah[m], bh[i][0], bh[i][1]);
        }
    }
    float* out = A.Y + ((size_t)blockIdx.x * A.K + k) * (MT * 16) * 192;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int col = 8 * (3 * w + i) + 2 * q;
            *reinterpret_cast<float2*>(out + (size_t)(16 * m + g) * 192 + col) = make_float2(acc[m][i][0], acc[m][i][1]);
            *reinterpret_cast<float2*>(out + (size_t)(16 * m + g + 8) * 192 + col) = make_float2(acc[m][i][2], acc[m][i][3]);
        }
}

template <typename Kern> static int set_smem(Kern kern, int bytes) {
    // the attribute is per device: set it on every launch (cheap) rather than caching per process
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? 0 : -1;
}

}  // namespace lgemm

static int sm_count() { static int n = 0; if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; } return n; }

// Y[T,K,N,n_out] = X . B_k ; x_has_tower = 0: X is [T,N,x_cols] (shared by the towers), 1: [T,K,N,x_cols].  Supported: (x_cols <= 40 | 48, n_out 192), (x_cols 192, n_out 48)
int launch_proj_rows(const float* X, int x_cols, int x_has_tower, const float* W, int w_trans, float* Y, int n_out, int T, int K, int N, cudaStream_t st) {
    using namespace lgemm;
    RowArgs a{}; a.X = X; a.x_cols = x_cols; a.x_t_stride = x_has_tower ? (long long)K * N : N; a.x_k_stride = x_has_tower ? N : 0;
    a.W = W; a.w_trans = w_trans; a.w_ld = w_trans ? x_cols : n_out; a.w_k_stride = (long long)x_cols * n_out; a.Y = Y; a.D = nullptr; a.T = T; a.K = K; a.N = N;
    const int tiles = T * ((N + 63) / 64);
    if (tiles <= 0) return 0;
    if (n_out == 192 && x_cols <= 40) {
        constexpr int B = sizeof(ProjSmem<5, 24>); if (set_smem(proj_rows_kernel<5, 24>, B)) return -2;
        proj_rows_kernel<5, 24><<<dim3(std::min(tiles, std::max(1, 2 * sm_count() / K)), K), THR, B, st>>>(a);
    } else if (n_out == 192 && x_cols == 48) {
        constexpr int B = sizeof(ProjSmem<6, 24>); if (set_smem(proj_rows_kernel<6, 24>, B)) return -2;
        proj_rows_kernel<6, 24><<<dim3(std::min(tiles, std::max(1, 2 * sm_count() / K)), K), THR, B, st>>>(a);
    } else if (n_out == 48 && x_cols == 192) {
        constexpr int B = sizeof(ProjSmem<24, 6>); if (set_smem(proj_rows_kernel<24, 6>, B)) return -2;
        proj_rows_kernel<24, 6><<<dim3(std::min(tiles, std::max(1, sm_count() / K)), K), THR, B, st>>>(a);
    } else return -1;
    return 0;
}
int gram_rows_ctas(int T, int N, int K) { const int tiles = T * ((N + 31) / 32); return std::max(1, std::min(tiles, 2 * sm_count() / std::max(K, 1))); }
// partial[gram_rows_ctas][K][48][192] += X^T D over this CTA's tiles; x_cols <= 48 (features past x_cols are zero rows)
int launch_gram_rows(const float* X, int x_cols, int x_has_tower, const float* D, float* partial, int T, int K, int N, cudaStream_t st) {
    using namespace lgemm;
    if (x_cols > 48 || x_cols <= 0) return -1;
    RowArgs a{}; a.X = X; a.x_cols = x_cols; a.x_t_stride = x_has_tower ? (long long)K * N : N; a.x_k_stride = x_has_tower ? N : 0;
    a.W = nullptr; a.Y = partial; a.D = D; a.T = T; a.K = K; a.N = N;
    constexpr int B = sizeof(GramSmem<3>); if (set_smem(gram_rows_kernel<3>, B)) return -2;
    gram_rows_kernel<3><<<dim3(gram_rows_ctas(T, N, K), K), THR, B, st>>>(a);
    return 0;
}

}  // namespace irrl
