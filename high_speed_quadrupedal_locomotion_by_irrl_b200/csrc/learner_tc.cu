// tcgen05 version of the learner's input projections  Y[t,k,n,0:192] = X[t,(k),n,:] . W_k  (ppo2.py:136-197 over run_bp_v5.py:151-166:
// x W_x of all T steps, both towers), the streaming product with the M = rows / N = 192 / K = 35 | 48 shape that fits a 128-lane TMEM
// accumulator exactly.  Same 3xTF32 arithmetic and the same operand layouts as the act kernel (policy_tc_kernels.cu, tc_common.cuh):
// UMMA canonical K-major no-swizzle tiles [k-step][hi|lo][16-byte k-chunk (2)][row][4 floats], tcgen05.mma.kind::tf32 M128 N192 K8
// (98 cycles each, 4x the rate of the warp-level MMAs of learner_gemm.cu), fp32 accumulators in tensor memory.
//
// One persistent CTA per SM (512 threads) walks 128-row tiles of one tower:
//   * raw rows arrive through the bulk-copy engine: ONE cp.async.bulk per tile (its rows are contiguous), completing on an mbarrier, two
//     tiles ahead.  Per-thread cp.async was measured slower: fence.proxy.async -- which every thread needs after writing the operand tile
//     -- waits for the thread's own outstanding asynchronous copies, i.e. one HBM round trip per tile.  (35-column rows need n_env % 4 == 0
//     for 16-byte alignment; otherwise the warp-level kernel of learner_gemm.cu serves the call.)
//   * every thread converts its fixed list of (row, 4 k) items into the hi / lo operand tile (row-per-lane writes: conflict-free; 48-column
//     rows are read as float4s in a rotated chunk order, conflict-free at the natural pitch of the copied block),
//   * one elected lane issues KS x 3 MMAs into TMEM buffer i % 2 and commits to an mbarrier,
//   * while they run, all threads drain the PREVIOUS tile's accumulator: tcgen05.ld (lane = row)
//     (24 columns per warp) -> padded staging tile -> coalesced float4 stores (a row-per-lane store straight from registers would cost 32 L1
//     tag look-ups per instruction).
// W is packed once per CTA (hi / lo split, K-major rows = output columns).  Rows past N in the last tile of a time step are never stored.
#include <algorithm>
#include <cstdlib>
#include "env_device.cuh"
#include "env_kernels.h"
#include "tc_common.cuh"

namespace irrl {
namespace ltc {
using namespace tc;

constexpr int THR = 512;                                        // 16 warps: the per-tile transform / staging work is issue- and latency-bound, not MMA-bound
constexpr int TM = 128, NG = 192;
constexpr int A_HALF = 2 * TM * 16, A_KSTEP = 2 * A_HALF;       // 4096 / 8192 bytes
constexpr int B_HALF = 2 * NG * 16, B_KSTEP = 2 * B_HALF;       // 6144 / 12288 bytes
constexpr int STAGE_PITCH = 100;                                // floats per row of the output staging tile (96 columns + 4): 16-byte groups 25 r mod 8 distinct
constexpr int TMEM_COLS = 512;                                  // 2 x 192 accumulator columns -> next power of two

__device__ __forceinline__ void cp16(void* s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(s)), "l"(g)); }
__device__ __forceinline__ void cp4(void* s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(s)), "l"(g)); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct ProjTcArgs {
    const float* X; int x_cols; long long x_t_stride, x_k_stride;      // row (t,k,n) of X starts at X + (t*x_t_stride + k*x_k_stride + n) * x_cols
    const float* W; long long w_k_stride;                              // W_k [x_cols][192]
    float* Y;                                                          // [T,K,N,192]
    int T, K, N;
};

template <int KS> struct Lay {                                         // KS k-steps of 8 (x_cols padded)
    static constexpr int RAW_PITCH = KS * 8 + 4;                       // floats; 16-byte groups (KS*2+1) r mod 8 distinct for row-per-lane float4 reads
    static constexpr int OFF_B = 0;
    static constexpr int OFF_A = OFF_B + KS * B_KSTEP;
    static constexpr int OFF_RAW = OFF_A + KS * A_KSTEP;
    static constexpr int OFF_STAGE = OFF_RAW + 2 * TM * RAW_PITCH * 4;
    static constexpr int OFF_BAR = OFF_STAGE + TM * STAGE_PITCH * 4;
    static constexpr int BYTES = OFF_BAR + 4 * 8 + 16;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}

// Everything a thread does per tile is the same list of (source, destination) pairs: the index arithmetic (divisions by the row length,
// operand-tile offsets) is done once before the tile loop and kept in registers.
template <int KS>
__global__ void __launch_bounds__(THR, 1) proj_rows_tc_kernel(const __grid_constant__ ProjTcArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using L = Lay<KS>;
    constexpr int RP = L::RAW_PITCH;
    static_assert(L::BYTES <= 232448, "shared memory");
    const int t_ = threadIdx.x, warp = t_ >> 5, lane = t_ & 31, k = blockIdx.y;
    unsigned char* sB = smem + L::OFF_B;
    unsigned char* sA = smem + L::OFF_A;
    float* raw = reinterpret_cast<float*>(smem + L::OFF_RAW);
    float* stage = reinterpret_cast<float*>(smem + L::OFF_STAGE);
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* raw_full = mma_done + 2;                                 // raw rows of a stage have landed (bulk-copy engine)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(raw_full + 2);
    constexpr bool VEC4 = (KS == 6);                                   // 48-column rows are read as float4s (rotated chunk order: conflict-free at the natural pitch), 35-column rows as scalars

    if (t_ == 0) { mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1); mbar_init(&raw_full[0], 1); mbar_init(&raw_full[1], 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
    // W_k -> B operand: row n of the tile = output column n, K-major, hi / lo split, zero rows past x_cols
    const float* Wk = A.W + (size_t)k * A.w_k_stride;
    for (int i = t_; i < NG * KS * 8; i += THR) {
        const int n = i % NG, kk = i / NG;
        const float v = kk < A.x_cols ? Wk[(size_t)kk * NG + n] : 0.f;
        const float hi = tf32_hi(v);
        *reinterpret_cast<float*>(sB + op_off(kk, n, NG)) = hi;
        *reinterpret_cast<float*>(sB + op_off(kk, n, NG) + B_HALF) = v - hi;
    }
    for (int i = t_; i < 2 * TM * RP; i += THR) raw[i] = 0.f;
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    const int tiles_per_t = (A.N + TM - 1) / TM, tiles = A.T * tiles_per_t;
    const int x_cols = A.x_cols;
    // ---- per-thread work lists
    constexpr int NIT = (TM * KS * 2 + THR - 1) / THR;                  // transform items (row, 4 k)
    uint32_t it_src[NIT], it_dst[NIT]; int it_nv[NIT];
#pragma unroll
    for (int j = 0; j < NIT; ++j) {
        // the raw tile is one contiguous block (rows x_cols apart).  Scalar reads (35 columns): 3 r mod 32 distinct over a warp.  float4 reads
        // (48 columns = 12 chunks): lane r takes chunk (jj + r / 2) mod 12, so that 8 consecutive rows hit 8 different 16-byte bank groups.
        const int i = t_ + j * THR, r = i & (TM - 1), jj = i >> 7, q = VEC4 ? (jj + (r >> 1)) % (KS * 2) : jj;
        it_src[j] = r * x_cols + 4 * q;
        it_nv[j] = min(4, max(0, x_cols - 4 * q));
        it_dst[j] = (jj < KS * 2) ? op_off(4 * q, r, TM) : 0xFFFFFFFFu;
    }
    constexpr int NCO = TM * 24 / THR;                                   // copy-out float4s per 96-column round
    uint32_t co_src[NCO], co_dst[NCO]; int co_row[NCO];
#pragma unroll
    for (int j = 0; j < NCO; ++j) { const int i = t_ + j * THR, r = i / 24, c4 = i - r * 24; co_row[j] = r; co_src[j] = r * STAGE_PITCH + 4 * c4; co_dst[j] = r * NG + 4 * c4; }

    // Raw rows come through the bulk-copy engine, not per-thread cp.async: fence.proxy.async (needed after the operand-tile writes) waits for
    // the issuing thread's outstanding asynchronous copies, which exposed one HBM round trip per tile.
    auto issue = [&](int tile, int buf) {
        const int t = tile / tiles_per_t, n0 = (tile - t * tiles_per_t) * TM, rows = min(TM, A.N - n0);
        const float* src = A.X + ((size_t)t * A.x_t_stride + (size_t)k * A.x_k_stride + n0) * x_cols;
        float* dst = raw + buf * TM * RP;
        if (t_ == 32) {                                                 // ONE copy per tile: per-row 192-byte copies were measured slower (3.1 vs 2.8 ms with cp.async)
            mbar_expect_tx(&raw_full[buf], rows * x_cols * 4);
            bulk_g2s(smem_u32(dst), src, rows * x_cols * 4, &raw_full[buf]);
        }
    };
    // accumulator of a finished tile: TMEM -> registers -> padded staging tile -> coalesced stores, 96 columns at a time
    auto drain = [&](int tile, int buf) {
        const int t = tile / tiles_per_t, n0 = (tile - t * tiles_per_t) * TM, rows = min(TM, A.N - n0);
        float* yb = A.Y + (((size_t)t * A.K + k) * A.N + n0) * NG;
        const int wq = warp & 3, sub = warp >> 2, row = 32 * wq + lane;            // 16 warps: TMEM lane quarter x 24-column block
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            uint32_t v[3][8];
#pragma unroll
            for (int i = 0; i < 3; ++i) tmem_ld8(tmem + ((uint32_t)(32 * wq) << 16) + buf * NG + 96 * h + 24 * sub + 8 * i, v[i]);
            tmem_ld_wait();
            float4* d = reinterpret_cast<float4*>(stage + row * STAGE_PITCH + 24 * sub);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                d[2 * i] = make_float4(__uint_as_float(v[i][0]), __uint_as_float(v[i][1]), __uint_as_float(v[i][2]), __uint_as_float(v[i][3]));
                d[2 * i + 1] = make_float4(__uint_as_float(v[i][4]), __uint_as_float(v[i][5]), __uint_as_float(v[i][6]), __uint_as_float(v[i][7]));
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < NCO; ++j)
                if (co_row[j] < rows) *reinterpret_cast<float4*>(yb + co_dst[j] + 96 * h) = *reinterpret_cast<const float4*>(stage + co_src[j]);
            __syncthreads();
        }
    };

    int tile = blockIdx.x, it = 0, prev_tile = -1;
    if (tile < tiles) issue(tile, 0);
    if (tile + (int)gridDim.x < tiles) issue(tile + gridDim.x, 1);
    for (; tile < tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&raw_full[b], (it >> 1) & 1);                         // the raw rows of this tile have landed
        if (it >= 1) { mbar_wait(&mma_done[b ^ 1], ((it - 1) >> 1) & 1); tc_fence_after(); }     // previous tile multiplied: the operand tile is free, its accumulator full
        const float* rw = raw + b * TM * RP;
#pragma unroll
        for (int j = 0; j < NIT; ++j) {                                 // item = (row, 4 k): row-per-lane reads and writes
            if (it_dst[j] != 0xFFFFFFFFu) {
                float v[4];
                if (VEC4) { const float4 x = *reinterpret_cast<const float4*>(rw + it_src[j]); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
                else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = e < it_nv[j] ? rw[it_src[j] + e] : 0.f;      // columns past x_cols are zero
                }
                store_hilo(sA, it_dst[j], A_HALF, v);
            }
        }
        fence_proxy_async();
        tc_fence_before(); __syncthreads(); tc_fence_after();
        {   // the raw buffer of this tile is free again: rows of the tile after next
            const int ahead = tile + 2 * (int)gridDim.x;
            if (ahead < tiles) issue(ahead, b);
        }
        if (warp == 0) {
            if (elect_one()) {
                const uint32_t idesc = make_idesc(TM, NG), d = tmem + b * NG;
#pragma unroll 1
                for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t ab = smem_u32(sA) + ks * A_KSTEP, bb = smem_u32(sB) + ks * B_KSTEP;
                    const uint64_t ah = make_desc(ab, TM * 16, 128), al = make_desc(ab + A_HALF, TM * 16, 128);
                    const uint64_t bh = make_desc(bb, NG * 16, 128), bl = make_desc(bb + B_HALF, NG * 16, 128);
                    mma_tf32(d, al, bh, idesc, ks > 0);
                    mma_tf32(d, ah, bl, idesc, 1);
                    mma_tf32(d, ah, bh, idesc, 1);
                }
                umma_commit(&mma_done[b]);
            }
            __syncwarp();
        }
        if (it >= 1) drain(prev_tile, b ^ 1);                           // overlaps the MMAs just issued
        prev_tile = tile;
    }
    if (it >= 1) {
        const int b = (it - 1) & 1;
        mbar_wait(&mma_done[b], ((it - 1) >> 1) & 1); tc_fence_after();
        drain(prev_tile, b);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, TMEM_COLS); }
}


// ---------------------------------------------------------------------------------------------------------------- projT_rows (tcgen05)
// Input gradient of a projection:  dX[t,k,n,0:48] = dz[t,k,n,0:192] . W_k^T  (W_k [48][192], so its rows are already K-major B rows).
// K = 192 does not fit one operand tile next to the rest, so a 128-row tile is multiplied as 4 column chunks of 48 (6 k-steps each)
// accumulating into the same 48 TMEM columns: the pipeline of proj_rows_tc_kernel runs over (tile, chunk) pairs, the accumulator is
// drained after every 4th pair.
constexpr int NT48 = 48;
constexpr int BT_HALF = 2 * NT48 * 16, BT_KSTEP = 2 * BT_HALF;         // 1536 / 3072 bytes
struct LayT {
    static constexpr int RAW_PITCH = 52;
    static constexpr int OFF_B = 0;                                    // 24 k-steps of W^T: 73,728 B
    static constexpr int OFF_A = OFF_B + 24 * BT_KSTEP;
    static constexpr int OFF_RAW = OFF_A + 6 * A_KSTEP;
    static constexpr int OFF_STAGE = OFF_RAW + 2 * TM * RAW_PITCH * 4;
    static constexpr int OFF_BAR = OFF_STAGE + TM * RAW_PITCH * 4;     // staging tile [128][52]
    static constexpr int BYTES = OFF_BAR + 4 * 8 + 16;
};
struct ProjTTcArgs { const float* D; const float* W; long long w_k_stride; float* Y; int T, K, N; };

__global__ void __launch_bounds__(THR, 1) projT_rows_tc_kernel(const __grid_constant__ ProjTTcArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using L = LayT;
    constexpr int RP = L::RAW_PITCH;
    const int t_ = threadIdx.x, warp = t_ >> 5, lane = t_ & 31, k = blockIdx.y;
    unsigned char* sB = smem + L::OFF_B;
    unsigned char* sA = smem + L::OFF_A;
    float* raw = reinterpret_cast<float*>(smem + L::OFF_RAW);
    float* stage = reinterpret_cast<float*>(smem + L::OFF_STAGE);
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* raw_full = mma_done + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(raw_full + 2);
    if (t_ == 0) { mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1); mbar_init(&raw_full[0], TM); mbar_init(&raw_full[1], TM); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(tmem_ptr, 128); tmem_relinquish(); }
    const float* Wk = A.W + (size_t)k * A.w_k_stride;                  // [48][192]
    for (int i = t_; i < NT48 * NG; i += THR) {                        // element (n, kk): chunk kk / 48, local k = kk % 48
        const int kk = i % NG, n = i / NG;
        const float v = Wk[(size_t)n * NG + kk];
        const float hi = tf32_hi(v);
        unsigned char* dst = sB + (kk / 48) * 6 * BT_KSTEP + op_off(kk % 48, n, NT48);
        *reinterpret_cast<float*>(dst) = hi; *reinterpret_cast<float*>(dst + BT_HALF) = v - hi;
    }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    const int tiles_per_t = (A.N + TM - 1) / TM, tiles = A.T * tiles_per_t;
    // virtual tile v = 4 * (index of this CTA's tile) + chunk
    uint32_t cp_src[6], cp_dst[6]; int cp_row[6];                        // copy chunks of the loader threads (256 threads x 6 = 128 rows x 12 chunks)
#pragma unroll
    for (int j = 0; j < 6; ++j) { const int i = (t_ & 255) + j * 256, r = i / 12, c = i - r * 12; cp_row[j] = r; cp_src[j] = r * NG + 4 * c; cp_dst[j] = r * RP + 4 * c; }
    uint32_t it_src[6], it_dst[6];                                       // transform items of the lower 256 threads
#pragma unroll
    for (int j = 0; j < 6; ++j) { const int i = (t_ & 255) + j * 256, r = i & (TM - 1), q = i >> 7; it_src[j] = r * RP + 4 * q; it_dst[j] = op_off(4 * q, r, TM); }
    uint32_t co_src[3], co_dst[3]; int co_row[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { const int i = t_ + j * THR, r = i / 12, c4 = i - r * 12; co_row[j] = r; co_src[j] = r * RP + 4 * c4; co_dst[j] = r * NT48 + 4 * c4; }
    const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x, vt_total = 4 * my_tiles;
    auto issue = [&](int vt, int buf) {
        const int tile = blockIdx.x + (vt >> 2) * gridDim.x, ch = vt & 3;
        const int t = tile / tiles_per_t, n0 = (tile - t * tiles_per_t) * TM, rows = A.N - n0;
        const float* src = A.D + (((size_t)t * A.K + k) * A.N + n0) * NG + 48 * ch;
        float* dst = raw + buf * TM * RP;
        // Column chunks of a [rows x 192] block are not contiguous, and per-row bulk copies are slow (6.0 vs 4.2 ms): cp.async it is, but issued by
        // the upper 8 warps only.  The lower 8 warps write the operand tile and execute fence.proxy.async, which waits for the issuing thread's own
        // outstanding asynchronous copies -- they have none.
        if (t_ >= 256) {
#pragma unroll
            for (int j = 0; j < 6; ++j) if (cp_row[j] < rows) cp16(dst + cp_dst[j], src + cp_src[j]);
            cp_commit();
        }
    };
    auto drain = [&](int tile, int buf) {
        const int t = tile / tiles_per_t, n0 = (tile - t * tiles_per_t) * TM, rows = min(TM, A.N - n0);
        float* yb = A.Y + (((size_t)t * A.K + k) * A.N + n0) * NT48;
        const int wq = warp & 3, sub = warp >> 2, row = 32 * wq + lane;
        if (sub < 2) {                                                  // 8 warps cover 128 lanes x 48 columns
            uint32_t v[3][8];
#pragma unroll
            for (int i = 0; i < 3; ++i) tmem_ld8(tmem + ((uint32_t)(32 * wq) << 16) + buf * NT48 + 24 * sub + 8 * i, v[i]);
            tmem_ld_wait();
            float4* d = reinterpret_cast<float4*>(stage + row * RP + 24 * sub);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                d[2 * i] = make_float4(__uint_as_float(v[i][0]), __uint_as_float(v[i][1]), __uint_as_float(v[i][2]), __uint_as_float(v[i][3]));
                d[2 * i + 1] = make_float4(__uint_as_float(v[i][4]), __uint_as_float(v[i][5]), __uint_as_float(v[i][6]), __uint_as_float(v[i][7]));
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (co_row[j] < rows) *reinterpret_cast<float4*>(yb + co_dst[j]) = *reinterpret_cast<const float4*>(stage + co_src[j]);
        __syncthreads();
    };
    if (vt_total > 0) issue(0, 0); else if (t_ >= 256) cp_commit();
    if (vt_total > 1) issue(1, 1); else if (t_ >= 256) cp_commit();
    for (int vt = 0; vt < vt_total; ++vt) {
        const int b = vt & 1, ch = vt & 3, tb = (vt >> 2) & 1;          // raw buffer, column chunk, TMEM buffer of this tile
        if (t_ >= 256) cp_wait<1>();                                    // loaders: every copy group but the newest has landed
        __syncthreads();
        if (vt >= 1) { mbar_wait(&mma_done[0], (vt - 1) & 1); tc_fence_after(); }      // one commit per virtual tile on mma_done[0]: operand tile free
        const float* rw = raw + b * TM * RP;
        if (t_ < 256) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float4 x = *reinterpret_cast<const float4*>(rw + it_src[j]);
                const float v[4] = {x.x, x.y, x.z, x.w};
                store_hilo(sA, it_dst[j], A_HALF, v);
            }
            fence_proxy_async();
        }
        tc_fence_before(); __syncthreads(); tc_fence_after();
        if (vt + 2 < vt_total) issue(vt + 2, b); else if (t_ >= 256) cp_commit();
        if (warp == 0) {
            if (elect_one()) {
                const uint32_t idesc = make_idesc(TM, NT48), d = tmem + tb * NT48;
#pragma unroll 1
                for (int ks = 0; ks < 6; ++ks) {
                    const uint32_t ab = smem_u32(sA) + ks * A_KSTEP, bb = smem_u32(sB) + (ch * 6 + ks) * BT_KSTEP;
                    const uint64_t ah = make_desc(ab, TM * 16, 128), al = make_desc(ab + A_HALF, TM * 16, 128);
                    const uint64_t bh = make_desc(bb, NT48 * 16, 128), bl = make_desc(bb + BT_HALF, NT48 * 16, 128);
                    mma_tf32(d, al, bh, idesc, (ch > 0 || ks > 0) ? 1u : 0u);
                    mma_tf32(d, ah, bl, idesc, 1);
                    mma_tf32(d, ah, bh, idesc, 1);
                }
                umma_commit(&mma_done[0]);
            }
            __syncwarp();
        }
        // the accumulator of the previous tile is complete once that tile's 4th chunk has been multiplied, i.e. at the wait above of vt = 4i:
        // drain it behind the MMAs just issued
        if (ch == 0 && vt >= 4) drain(blockIdx.x + ((vt >> 2) - 1) * gridDim.x, tb ^ 1);
    }
    if (vt_total > 0) {
        mbar_wait(&mma_done[0], (vt_total - 1) & 1); tc_fence_after();
        drain(blockIdx.x + (my_tiles - 1) * gridDim.x, (my_tiles - 1) & 1);
    }
    cp_wait<0>();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 128); }
}

// ---------------------------------------------------------------------------------------------------------------- gram2_rows (tcgen05)
// Both weight gradients of one LSTM layer in ONE pass over dz:  G[0:48] = sum_rows x^T dz (dW_x),  G[48:96] = sum_rows hm^T dz (dW_h),
// hm(t) = h(t-1) * keep(t) formed here from the layer's OUTPUT sequence (the forward kernel does not store a masked copy),
// rows = (t, n) of tower k (~6 M).  MMA view: D[128 x 192] += A[128 x 8] . B[8 x 192] with M = feature (96 used), N = gate column, K = ROWS,
// so both operands must be K-major over rows: the transform transposes while it splits -- an item is (feature or column, 4 consecutive
// rows): four strided scalar reads (lanes along the feature: conflict-free), hi / lo split, one float4 store each into the operand tile.
// The fp32 accumulator stays in tensor memory over all tiles of the CTA; one partial [128 x 192] per CTA is written at the end (summed
// afterwards in a fixed order: deterministic).  Raw rows stream through a double-buffered cp.async stage, 32 rows (4 k-steps) per tile.
struct GramTcArgs {
    const float* X; int x_cols; long long x_t_stride, x_k_stride;      // [T,(K),N,x_cols], x_cols <= 48
    const float* Hs; const float* h0; const float* keep;               // masked input state of step t = Hs[t-1] * keep[t] (h0 * keep[0] at t = 0): Hs [T,K,N,48], h0 [K,N,48], keep [T,N]
    const float* D;                                                    // [T,K,N,192]
    float* P;                                                          // partial sums [gridDim.x][K][128][192]
    int T, K, N;
};
constexpr int GR = 24, GKS = GR / 8;                                   // rows per tile = 3 k-steps: small enough for TWO operand-tile slots and THREE raw stages
constexpr int G_SLOT = GKS * (A_KSTEP + B_KSTEP);                      // operand tiles of one slot: A (128 x 24) then B (192 x 24), hi / lo each
constexpr int G_OFF_OP = 0, G_OFF_RAW = G_OFF_OP + 2 * G_SLOT;
constexpr int G_RAW_STAGE = GR * (48 + 48 + NG + 4) * 4;               // [24][48] x | [24][48] h(t-1) | [24][192] dz | [24] keep(t) (+ pad to 16 bytes per 4 rows)
constexpr int G_NRAW = 3;
constexpr int G_OFF_BAR = G_OFF_RAW + G_NRAW * G_RAW_STAGE;
constexpr int G_BYTES = G_OFF_BAR + (2 + G_NRAW) * 8 + 16;
static_assert(G_BYTES <= 232448, "shared memory");

// Pipeline: the transform of tile i+1 (into the other operand slot) runs while the MMAs of tile i execute; a slot is reused when the
// MMAs of tile i-2 have committed; raw rows are three tiles ahead.
__global__ void __launch_bounds__(THR, 1) gram2_rows_tc_kernel(const __grid_constant__ GramTcArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t_ = threadIdx.x, warp = t_ >> 5, lane = t_ & 31, k = blockIdx.y;
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + G_OFF_BAR);            // [2]: one per operand slot
    uint64_t* raw_full = mma_done + 2;                                             // [G_NRAW]: bulk copies of a raw stage have landed
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(raw_full + G_NRAW);
    if (t_ == 0) { mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1); for (int i = 0; i < G_NRAW; ++i) mbar_init(&raw_full[i], 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(tmem_ptr, 256); tmem_relinquish(); }
    for (int i = t_; i < (G_OFF_BAR - G_OFF_OP) / 4; i += THR) reinterpret_cast<float*>(smem)[i] = 0.f;     // operand rows 96..127, feature columns past x_cols: zero for good
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    const int tiles_per_t = (A.N + GR - 1) / GR, tiles = A.T * tiles_per_t;
    const int x_cols = A.x_cols;
    uint32_t a_src[2], a_dst[2], b_src[3], b_dst[3]; int a_pitch[2], a_keep[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = t_ + j * THR, m = i % 96, c = i / 96;
        a_pitch[j] = m < 48 ? x_cols : 48; a_keep[j] = m < 48 ? -1 : 4 * c;
        a_src[j] = m < 48 ? m + 4 * c * x_cols : GR * 48 + (m - 48) + 4 * c * 48;
        a_dst[j] = (i < 96 * (GR / 4) && (m >= 48 || m < x_cols)) ? op_off(4 * c, m, TM) : 0xFFFFFFFFu;      // features past x_cols: their operand rows stay zero
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) { const int i = t_ + j * THR, n = i % NG, c = i / NG; b_src[j] = GR * 96 + 4 * c * NG + n; b_dst[j] = i < NG * (GR / 4) ? GKS * A_KSTEP + op_off(4 * c, n, NG) : 0xFFFFFFFFu; }

    // Raw rows come through the bulk-copy engine (three contiguous blocks per tile, completion on an mbarrier), NOT through per-thread cp.async:
    // fence.proxy.async, which every thread needs after writing the operand tile, waits for the thread's own outstanding asynchronous copies,
    // i.e. for a full HBM round trip per tile (measured: ~3 k cycles per tile whatever its size).
    auto issue = [&](int tile, int buf) {
        const int t = tile / tiles_per_t, n0 = (tile - t * tiles_per_t) * GR, rows = min(GR, A.N - n0);
        float* rx = reinterpret_cast<float*>(smem + G_OFF_RAW + buf * G_RAW_STAGE);
        if (rows < GR) {                                                // ragged last tile of a time step: the rows past N must be zero
            for (int i = t_; i < (GR - rows) * x_cols; i += THR) rx[rows * x_cols + i] = 0.f;
            for (int i = t_; i < (GR - rows) * 48; i += THR) rx[GR * 48 + rows * 48 + i] = 0.f;
            for (int i = t_; i < (GR - rows) * NG; i += THR) rx[GR * 96 + rows * NG + i] = 0.f;
        }
        if (t_ == 32) {
            const float* xs = A.X + ((size_t)t * A.x_t_stride + (size_t)k * A.x_k_stride + n0) * x_cols;
            const float* hs = t > 0 ? A.Hs + (((size_t)(t - 1) * A.K + k) * A.N + n0) * 48 : A.h0 + ((size_t)k * A.N + n0) * 48;
            const float* ks = A.keep + (size_t)t * A.N + n0;
            const float* ds = A.D + (((size_t)t * A.K + k) * A.N + n0) * NG;
            const uint32_t bx = rows * x_cols * 4, bh = rows * 48 * 4, bd = rows * NG * 4;
            mbar_expect_tx(&raw_full[buf], bx + bh + bd + rows * 4);
            bulk_g2s(smem_u32(rx), xs, bx, &raw_full[buf]);
            bulk_g2s(smem_u32(rx + GR * 48), hs, bh, &raw_full[buf]);
            bulk_g2s(smem_u32(rx + GR * 96), ds, bd, &raw_full[buf]);
            bulk_g2s(smem_u32(rx + GR * 288), ks, rows * 4, &raw_full[buf]);
        }
    };
    int tile = blockIdx.x, it = 0;
#pragma unroll
    for (int j = 0; j < G_NRAW; ++j) if (tile + j * (int)gridDim.x < tiles) issue(tile + j * gridDim.x, j);
    __syncthreads();                                                    // the zero rows a ragged first tile got from other threads (later refills sit behind the per-tile barriers)
    for (; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it & 1, rb = it % G_NRAW;
        mbar_wait(&raw_full[rb], (it / G_NRAW) & 1);                    // the raw rows of this tile have landed
        if (it >= 2) { mbar_wait(&mma_done[s], ((it - 2) >> 1) & 1); tc_fence_after(); }     // the MMAs of tile it-2 have read this operand slot
        unsigned char* sA = smem + G_OFF_OP + s * G_SLOT;
        const float* rx = reinterpret_cast<const float*>(smem + G_OFF_RAW + rb * G_RAW_STAGE);
#pragma unroll
        for (int j = 0; j < 2; ++j) {                                   // A: (feature m, rows 4c .. 4c+3); features 0..47 = x, 48..95 = hm
            if (a_dst[j] != 0xFFFFFFFFu) {
                const float* col = rx + a_src[j];
                const int ap = a_pitch[j];
                float v[4] = {col[0], col[ap], col[2 * ap], col[3 * ap]};
                if (a_keep[j] >= 0) {                                   // h(t-1) rows: SB lstm() masks the state fed into step t (h *= 1 - m)
                    const float4 kp = *reinterpret_cast<const float4*>(rx + GR * 288 + a_keep[j]);
                    v[0] *= kp.x; v[1] *= kp.y; v[2] *= kp.z; v[3] *= kp.w;
                }
                store_hilo(sA, a_dst[j], A_HALF, v);
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {                                   // B: (gate column n, rows 4c .. 4c+3)
            if (b_dst[j] != 0xFFFFFFFFu) {
                const float* col = rx + b_src[j];
                const float v[4] = {col[0], col[NG], col[2 * NG], col[3 * NG]};
                store_hilo(sA, b_dst[j], B_HALF, v);
            }
        }
        fence_proxy_async();
        tc_fence_before(); __syncthreads(); tc_fence_after();
        {
            const int ahead = tile + G_NRAW * (int)gridDim.x;
            if (ahead < tiles) issue(ahead, rb);
        }
        if (warp == 0) {
            if (elect_one()) {
                const uint32_t idesc = make_idesc(TM, NG);
#pragma unroll 1
                for (int ks = 0; ks < GKS; ++ks) {
                    const uint32_t ab = smem_u32(sA) + ks * A_KSTEP, bb = smem_u32(sA) + GKS * A_KSTEP + ks * B_KSTEP;
                    const uint64_t ah = make_desc(ab, TM * 16, 128), al = make_desc(ab + A_HALF, TM * 16, 128);
                    const uint64_t bh = make_desc(bb, NG * 16, 128), bl = make_desc(bb + B_HALF, NG * 16, 128);
                    mma_tf32(tmem, al, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    mma_tf32(tmem, ah, bl, idesc, 1);
                    mma_tf32(tmem, ah, bh, idesc, 1);
                }
                umma_commit(&mma_done[s]);
            }
            __syncwarp();
        }
    }
    // all MMAs complete when the last commit of each slot has arrived (commits are in issue order)
    if (it >= 2) { const int j = it - 2; mbar_wait(&mma_done[j & 1], (j >> 1) & 1); }
    if (it >= 1) { const int j = it - 1; mbar_wait(&mma_done[j & 1], (j >> 1) & 1); }
    tc_fence_after();
    if (warp < 4) {                                                     // lane = feature row m: 96 of them are real
        const int m = 32 * warp + lane;
        float* out = A.P + (((size_t)blockIdx.x * A.K + k) * TM + m) * NG;
#pragma unroll 1
        for (int i = 0; i < NG / 16; ++i) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + 16 * i, v);
            tmem_ld_wait();
            if (it >= 1 && m < 96) {
#pragma unroll
                for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(out + 16 * i)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 256); }
}

}  // namespace ltc

static int sm_count_tc() { static int n = 0; if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; } return n; }

// 0 (default) = tcgen05 kernel of this file for the forward projections, 1 = warp-level MMA kernels of learner_gemm.cu (A/B, regression reference)
static int g_proj_path = [] { const char* e = getenv("IRRL_PROJ_PATH"); return (e && e[0] == 'm') ? 1 : 0; }();
int proj_rows_set_path(int path) { const int prev = g_proj_path; if (path == 0 || path == 1) g_proj_path = path; return prev; }

// Y[T,K,N,192] = X . W_k on tcgen05; x_cols <= 40 (shared observation or per-tower) or 48.  Returns -1 for other shapes.
int launch_proj_rows_tc(const float* X, int x_cols, int x_has_tower, const float* W, float* Y, int T, int K, int N, cudaStream_t st) {
    using namespace ltc;
    if (g_proj_path == 1) return -1;
    ProjTcArgs a{}; a.X = X; a.x_cols = x_cols; a.x_t_stride = x_has_tower ? (long long)K * N : N; a.x_k_stride = x_has_tower ? N : 0;
    a.W = W; a.w_k_stride = (long long)x_cols * NG; a.Y = Y; a.T = T; a.K = K; a.N = N;
    const int tiles = T * ((N + TM - 1) / TM);
    if (tiles <= 0) return 0;
    const dim3 grid(std::min(tiles, std::max(1, sm_count_tc() / K)), K);
    if (reinterpret_cast<uintptr_t>(X) & 15) return -1;                 // bulk copies want 16-byte aligned sources
    if (x_cols <= 40 && x_cols > 32) {
        if ((x_cols & 3) && (N & 3)) return -1;                          // whole-block copies: every tile must start on and span a multiple of 16 bytes
        if (cudaFuncSetAttribute(proj_rows_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<5>::BYTES) != cudaSuccess) return -2;
        proj_rows_tc_kernel<5><<<grid, THR, Lay<5>::BYTES, st>>>(a);
    } else if (x_cols == 48) {
        if (cudaFuncSetAttribute(proj_rows_tc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<6>::BYTES) != cudaSuccess) return -2;
        proj_rows_tc_kernel<6><<<grid, THR, Lay<6>::BYTES, st>>>(a);
    } else return -1;
    return 0;
}

}  // namespace irrl

namespace irrl {
// dX[T,K,N,48] = D[T,K,N,192] . W_k^T (W [K,48,192]) on tcgen05; -1 when the warp-level path is selected
int launch_projT_rows_tc(const float* D, const float* W, float* Y, int T, int K, int N, cudaStream_t st) {
    using namespace ltc;
    if (g_proj_path == 1 || (reinterpret_cast<uintptr_t>(D) & 15)) return -1;
    ProjTTcArgs a{}; a.D = D; a.W = W; a.w_k_stride = 48 * 192; a.Y = Y; a.T = T; a.K = K; a.N = N;
    const int tiles = T * ((N + TM - 1) / TM);
    if (tiles <= 0) return 0;
    if (cudaFuncSetAttribute(projT_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LayT::BYTES) != cudaSuccess) return -2;
    projT_rows_tc_kernel<<<dim3(std::min(tiles, std::max(1, sm_count_tc() / K)), K), THR, LayT::BYTES, st>>>(a);
    return 0;
}
int gram2_rows_tc_ctas(int T, int N, int K) { const int tiles = T * ((N + ltc::GR - 1) / ltc::GR); return std::max(1, std::min(tiles, sm_count_tc() / std::max(K, 1))); }
// partial[gram2_rows_tc_ctas][K][128][192]: rows 0..x_cols-1 = sum x^T dz, rows 48..95 = sum hm^T dz (rows 96..127 are not written)
int launch_gram2_rows_tc(const float* X, int x_cols, int x_has_tower, const float* Hs, const float* h0, const float* keep, const float* D, float* partial, int T, int K, int N,
                         cudaStream_t st) {
    using namespace ltc;
    if (x_cols <= 0 || x_cols > 48) return -1;
    // the bulk-copy engine wants 16-byte aligned blocks of a multiple of 16 bytes: true for every tile when N is a multiple of 4
    if ((N & 3) || ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Hs) | reinterpret_cast<uintptr_t>(h0) | reinterpret_cast<uintptr_t>(keep) | reinterpret_cast<uintptr_t>(D)) & 15)) return -3;
    GramTcArgs a{}; a.X = X; a.x_cols = x_cols; a.x_t_stride = x_has_tower ? (long long)K * N : N; a.x_k_stride = x_has_tower ? N : 0;
    a.Hs = Hs; a.h0 = h0; a.keep = keep; a.D = D; a.P = partial; a.T = T; a.K = K; a.N = N;
    if (cudaFuncSetAttribute(gram2_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_BYTES) != cudaSuccess) return -2;
    gram2_rows_tc_kernel<<<dim3(gram2_rows_tc_ctas(T, N, K), K), THR, G_BYTES, st>>>(a);
    return 0;
}
}  // namespace irrl
