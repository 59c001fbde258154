// Tensor-core (tcgen05 + TMEM + bulk-copy) version of the fused two-tower 2x48 LSTM policy act() for sm_100a.
//
// Same contract as lstm_act_kernel (policy_kernels.cu; CustomLSTMPolicy.step run_bp_v5.py:178-185 over the graph
// run_bp_v5.py:143-176): both layers of one tower, the pi / V head, the Gaussian sample and neglogp in one launch.
//
// Mapping.  A CTA owns 128 environments (the MMA M dimension, one TMEM lane per environment) and one tower
// (blockIdx.y).  Every layer is one GEMM  gates[128 x 192] = [x ; h][128 x K] * W^T[K x 192]  issued by a single
// thread as tcgen05.mma.kind::tf32 instructions (M128 N192 K8) with the fp32 accumulator in tensor memory.
//   * fp32 accuracy on tf32 hardware: both operands are split x = hi + lo (hi = the 19 leading bits, exactly a tf32
//     number; lo = x - hi) and three products are accumulated, lo*hi + hi*lo + hi*hi ("3xTF32"); the dropped lo*lo
//     term is 2^-22 relative.  Agreement with the fp32 FMA kernel and the float64 oracle stays inside the 2e-5 bar.
//   * operands live in shared memory in the UMMA canonical K-major no-swizzle layout
//     [k-step of 8][hi|lo][16-byte k-chunk (2)][row][4 floats]  (core matrix = 8 rows x 16 B, SBO = 128 B,
//     LBO = rows * 16 B), so a thread that owns one environment row writes its 4 new hidden units as one float4.
//   * the weights are stored once (irrl_policy_set_params) in exactly that layout, hi and lo pre-split, and stream
//     L2 -> shared memory through a 6-stage ring of 12 KB cp.async.bulk copies completing on mbarriers (TMA engine);
//     tcgen05.commit releases a stage as soon as its three MMAs have read it.
//   * epilogue: 8 warps; warp w reads TMEM lanes 32 (w%4).., gate columns 96 (w/4).. with tcgen05.ld.32x32b.x16 --
//     columns are gate-interleaved (4 unit + gate), so 16 columns = 4 hidden units x (i,f,o,g): the cell update is done
//     in registers, c/h go to HBM and h (hi/lo) straight back into the A-operand tile of the next GEMM.
//   * the heads are a third GEMM (N = 16) on h of the top layer; lanes then sample / write their own environment.
// Roles: warps 0-7 stage + epilogue, warp 8 lane 0 = bulk-copy producer, warp 9 lane 0 = MMA issuer (warp 9 owns TMEM).
#include <cstring>
#include "env_device.cuh"
#include "env_kernels.h"

namespace irrl {
namespace tc {

constexpr int TM = 128;                          // environments per CTA = MMA M
constexpr int NG = 192;                          // gate columns = MMA N
constexpr int NHEAD = 16;                        // head columns (12 means / 1 value, zero padded) = MMA N of the head GEMM
constexpr int KS1 = 11, KS2 = 12, KSH = 6;       // k-steps (8 floats) of layer 0 [obs 35 -> 40 ; h 48], layer 1 [h_new 48 ; h 48], head [h 48]
constexpr int A_HALF = 2 * TM * 16;              // bytes of the hi (or lo) part of one A k-step : [chunk 2][row 128][16 B]
constexpr int A_KSTEP = 2 * A_HALF;              // 8192
constexpr int B_HALF = 2 * NG * 16;              // 6144
constexpr int B_KSTEP = 2 * B_HALF;              // 12288
constexpr int H_HALF = 2 * NHEAD * 16;           // 512
constexpr int H_KSTEP = 2 * H_HALF;              // 1024
constexpr int NST = 6;                           // ring stages
constexpr int NWORK = 512;                       // staging / epilogue threads (16 warps: TMEM lane quarter w%4, 48-column block w/4)
constexpr int NTHR = NWORK + 64;
constexpr int TMEM_COLS = 512;                   // 192 (layer 0) + 192 (layer 1) + 16 (head) -> next power of two
constexpr int OFF_A1 = 0;
constexpr int OFF_A2 = OFF_A1 + KS1 * A_KSTEP;   // 90112 : h(t-1) of layer 1
constexpr int OFF_RING = OFF_A2 + 6 * A_KSTEP;   // 139264
constexpr int OFF_BIAS = OFF_RING + NST * B_KSTEP;   // 212992 : 2 x 192 gate biases, 16 head biases, 16 logstd
constexpr int OFF_BAR = OFF_BIAS + (2 * NG + 32) * 4;
constexpr int SMEM_BYTES = OFF_BAR + (2 * NST + 4) * 8 + 16;
static_assert((KS1 + KS2) * B_KSTEP + KSH * H_KSTEP == TC_BLOB_BYTES, "blob size");
static_assert(SMEM_BYTES <= 232448, "shared memory");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch failure reported to the caller) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    if (mbar_try_wait(b, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(b, parity)) { if (clock64() - t0 > 4000000000ll) __trap(); }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t cols) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory"); }
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32 (bits 4-5 = 1), A and B tf32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at 17, M >> 4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { float e = __expf(-2.0f * fabsf(x)); float r = __fdividef(1.0f - e, 1.0f + e); return copysignf(r, x); }

// byte offset of element (row, k) inside an operand tile of `rows` rows (hi part; lo part follows at +2*rows*16)
__device__ __host__ __forceinline__ uint32_t op_off(int k, int row, int rows) { return (uint32_t)((k >> 3) * (4 * rows * 16) + ((k >> 2) & 1) * (rows * 16) + row * 16 + (k & 3) * 4); }
// nearest tf32 number (10 explicit mantissa bits): x = hi + lo with |lo| <= 2^-12 |x|, both exact in fp32
__device__ __host__ __forceinline__ float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
    uint32_t b; memcpy(&b, &x, 4); b = (b + 0x1000u) & 0xFFFFE000u; float h; memcpy(&h, &b, 4); return h;
#endif
}
// split 4 consecutive k of one row into tf32-exact hi and the remainder lo, one float4 store each
__device__ __forceinline__ void store_hilo(unsigned char* tile, uint32_t off, uint32_t lo_off, const float (&v)[4]) {
    float4 hi, lo;
    hi.x = tf32_hi(v[0]); lo.x = v[0] - hi.x;
    hi.y = tf32_hi(v[1]); lo.y = v[1] - hi.y;
    hi.z = tf32_hi(v[2]); lo.z = v[2] - hi.z;
    hi.w = tf32_hi(v[3]); lo.w = v[3] - hi.w;
    *reinterpret_cast<float4*>(tile + off) = hi;
    *reinterpret_cast<float4*>(tile + off + lo_off) = lo;
}

// optional timeline of CTA (0,0) in SM clocks (irrl_tc_timeline): a cheap way to see which hand-off paces the kernel
__device__ long long g_timeline[16];
__device__ int g_timeline_on = 0;
#define TC_MARK(slot) do { if (g_timeline_on && blockIdx.x == 0 && blockIdx.y == 0) g_timeline[slot] = clock64(); } while (0)

__global__ void __launch_bounds__(NTHR, 1) lstm_act_tc_kernel(const __grid_constant__ ActArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tower = blockIdx.y, e0 = blockIdx.x * TM;
    unsigned char* sA1 = smem + OFF_A1;
    unsigned char* sA2 = smem + OFF_A2;
    float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);      // [2][192] gate biases, [16] head bias, [16] logstd
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty = full + NST;
    uint64_t* a_ready = empty + NST;
    uint64_t* d_full = a_ready + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_full + 3);

    if (t == NWORK) {
        for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(a_ready, NWORK);
        for (int i = 0; i < 3; ++i) mbar_init(&d_full[i], 1);
        fence_mbar_init();
    }
    if (warp == NWORK / 32 + 1) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
    if (t < NWORK) {
        for (int i = t; i < 2 * NG; i += NWORK) sbias[i] = A.W.bperm[tower * 2 + i / NG][i % NG];
        if (t < 16) sbias[2 * NG + t] = tower == 0 ? (t < ACT_DIM ? A.W.pi_b[t] : 0.f) : (t == 0 ? A.W.vf_b[0] : 0.f);
        else if (t < 32) sbias[2 * NG + t] = (t - 16) < ACT_DIM ? A.W.logstd[t - 16] : 0.f;
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (t == 0) TC_MARK(0);

    if (t < NWORK) {
        // ------------------------------------------------------------ stage the A operands (hi/lo split)
        // every global load of a thread is issued before its first use, so one HBM latency is exposed, not a chain of them
        const int nvalid = min(TM, A.N - e0) * OB_DIM;                // this tile's observations are one contiguous block
        float* stage = reinterpret_cast<float*>(sA2);                 // layer-1 tile doubles as the obs staging buffer until h1 is stored
        float ov[9]; float4 h0[3], h1[3]; float kp[3];
#pragma unroll
        for (int it = 0; it < 9; ++it) { const int i = t + it * NWORK; ov[it] = (i < nvalid) ? __ldg(A.obs + (size_t)e0 * OB_DIM + i) : 0.f; }
#pragma unroll
        for (int it = 0; it < 3; ++it) {                             // h(t-1) of both layers: item = (row, 4 units)
            const int idx = t + it * NWORK, r = idx & (TM - 1), q = idx >> 7, env = min(e0 + r, A.N - 1);
            const float* st = A.state + (size_t)env * LSTM_STATE + tower * 4 * LSTM_H;
            h0[it] = *reinterpret_cast<const float4*>(st + LSTM_H + 4 * q);
            h1[it] = *reinterpret_cast<const float4*>(st + 3 * LSTM_H + 4 * q);
            kp[it] = (A.done && A.done[env]) ? 0.f : 1.f;
        }
        const int wq = warp & 3, part = warp >> 2, row = 32 * wq + lane, env = e0 + row, envc = min(env, A.N - 1);
        const bool valid = env < A.N;
        const float keep = (A.done && A.done[envc]) ? 0.f : 1.f;
        float4 cprev[2][3];                                           // c(t-1) of this lane's 12 units, both layers (needed after the first GEMM)
#pragma unroll
        for (int l = 0; l < 2; ++l)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                cprev[l][i] = *reinterpret_cast<const float4*>(A.state + (size_t)envc * LSTM_STATE + tower * 4 * LSTM_H + l * 2 * LSTM_H + 12 * part + 4 * i);
#pragma unroll
        for (int it = 0; it < 9; ++it) {
            const int i = t + it * NWORK;
            if (i < TM * OB_DIM) stage[i] = ov[it];
            if (tower == 0 && A.obs_store && i < nvalid) A.obs_store[(size_t)e0 * OB_DIM + i] = ov[it];      // mb_obs (ppo2.py:522)
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");
#pragma unroll
        for (int it = 0; it < 3; ++it) {                             // observations: item = (row, 4 k), zero padded 35 -> 40
            const int idx = t + it * NWORK, r = idx & (TM - 1), q = idx >> 7;
            if (q < 10) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (4 * q + j < OB_DIM) ? stage[r * OB_DIM + 4 * q + j] : 0.f;
                store_hilo(sA1, op_off(4 * q, r, TM), A_HALF, v);
            }
        }
#pragma unroll
        for (int it = 0; it < 3; ++it) {                             // masked (SB lstm(): h *= 1-m)
            const int idx = t + it * NWORK, r = idx & (TM - 1), q = idx >> 7;
            const float v0[4] = {h0[it].x * kp[it], h0[it].y * kp[it], h0[it].z * kp[it], h0[it].w * kp[it]};
            store_hilo(sA1, op_off(40 + 4 * q, r, TM), A_HALF, v0);
        }
        if (tower == 0 && A.done_store && t < TM && e0 + t < A.N) A.done_store[e0 + t] = A.done ? A.done[e0 + t] : 0;   // mb_dones (ppo2.py:526)
        fence_proxy_async();
        mbar_arrive(a_ready);
        if (t == 0) TC_MARK(1);
        asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");      // everyone is done reading the staging buffer
#pragma unroll
        for (int it = 0; it < 3; ++it) {                             // h(t-1) of layer 1 (consumed by the second GEMM only)
            const int idx = t + it * NWORK, r = idx & (TM - 1), q = idx >> 7;
            const float v1[4] = {h1[it].x * kp[it], h1[it].y * kp[it], h1[it].z * kp[it], h1[it].w * kp[it]};
            store_hilo(sA2, op_off(4 * q, r, TM), A_HALF, v1);
        }

        // ------------------------------------------------------------ cell updates straight out of tensor memory
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            mbar_wait(&d_full[l], 0);
            tc_fence_after();
            if (t == 0) TC_MARK(5 + 4 * l);
            float* st = A.state + (size_t)envc * LSTM_STATE + tower * 4 * LSTM_H + l * 2 * LSTM_H;
            const uint32_t taddr = tmem + ((uint32_t)(32 * wq) << 16) + l * NG + 48 * part;
            const float* bl = sbias + l * NG + 48 * part;
            uint32_t v[3][16];
#pragma unroll
            for (int i = 0; i < 3; ++i) tmem_ld16(taddr + 16 * i, v[i]);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int u = 12 * part + 4 * i;
                const float cold[4] = {cprev[l][i].x, cprev[l][i].y, cprev[l][i].z, cprev[l][i].w};
                float cn[4], hn[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {      // gate order i,f,o,g (CustomerLstmNN.py:119-126); reciprocals shared pairwise
                    const float4 b4 = *reinterpret_cast<const float4*>(bl + 16 * i + 4 * j);
                    const float zi = __uint_as_float(v[i][4 * j + 0]) + b4.x, zf = __uint_as_float(v[i][4 * j + 1]) + b4.y;
                    const float zo = __uint_as_float(v[i][4 * j + 2]) + b4.z, zg = __uint_as_float(v[i][4 * j + 3]) + b4.w;
                    const float ei = __expf(-zi), ef = __expf(-zf), eo = __expf(-zo), eg = __expf(-2.0f * fabsf(zg));
                    const float ig_gg = copysignf(__fdividef(1.0f - eg, (1.0f + ei) * (1.0f + eg)), zg);       // sigmoid(zi) * tanh(zg)
                    const float fg = __fdividef(1.0f, 1.0f + ef);
                    cn[j] = fmaf(fg, cold[j] * keep, ig_gg);
                    const float ec = __expf(-2.0f * fabsf(cn[j]));
                    hn[j] = copysignf(__fdividef(1.0f - ec, (1.0f + eo) * (1.0f + ec)), cn[j]);                 // sigmoid(zo) * tanh(c)
                }
                if (valid) {
                    *reinterpret_cast<float4*>(st + u) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    *reinterpret_cast<float4*>(st + LSTM_H + u) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                }
                store_hilo(sA1, op_off(u, row, TM), A_HALF, hn);      // A operand of the next GEMM (k = unit index)
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(a_ready);
            if (t == 0) TC_MARK(6 + 4 * l);
        }
        // ------------------------------------------------------------ heads: lanes of warps 0-3 own one environment each
        mbar_wait(&d_full[2], 0);
        tc_fence_after();
        if (t == 0) TC_MARK(12);
        if (part == 0) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(32 * wq) << 16) + 2 * NG, v);
            tmem_ld_wait();
            const float* hb = sbias + 2 * NG;
            if (tower == 0) {
                float act[ACT_DIM], mean[ACT_DIM];
                float nlp[ACT_DIM];
#pragma unroll
                for (int g3 = 0; g3 < 3; ++g3) {
                    float g[4] = {0.f, 0.f, 0.f, 0.f};
                    if (!A.deterministic) gauss4(A.seed, (uint32_t)envc + A.env_offset, A.tick, P_POLICY_EPS + g3, g);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int a = 4 * g3 + j;
                        const float m = __uint_as_float(v[a]) + hb[a], ls = hb[16 + a], sd = expf(ls);
                        const float x = fmaf(sd, g[j], m), z = (x - m) / sd;
                        mean[a] = m; act[a] = x; nlp[a] = 0.5f * z * z + ls;
                    }
                }
                float acc = 0.5f * 1.8378770664093453f * ACT_DIM;     // 0.5 * ln(2 pi) * 12
#pragma unroll
                for (int a = 0; a < ACT_DIM; ++a) acc += nlp[a];
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        reinterpret_cast<float4*>(A.action + (size_t)env * ACT_DIM)[q] = make_float4(act[4 * q], act[4 * q + 1], act[4 * q + 2], act[4 * q + 3]);
                        if (A.clipped)
                            reinterpret_cast<float4*>(A.clipped + (size_t)env * ACT_DIM)[q] =
                                make_float4(fminf(fmaxf(act[4 * q], -1.f), 1.f), fminf(fmaxf(act[4 * q + 1], -1.f), 1.f), fminf(fmaxf(act[4 * q + 2], -1.f), 1.f),
                                            fminf(fmaxf(act[4 * q + 3], -1.f), 1.f));                                   // ppo2.py:529-531
                        if (A.mean) reinterpret_cast<float4*>(A.mean + (size_t)env * ACT_DIM)[q] = make_float4(mean[4 * q], mean[4 * q + 1], mean[4 * q + 2], mean[4 * q + 3]);
                    }
                    A.neglogp[env] = acc;
                }
            } else if (valid) {
                A.value[env] = __uint_as_float(v[0]) + hb[0];
            }
        }
    } else if (t == NWORK) {
        // ------------------------------------------------------------ weight producer: bulk copies through the ring
        const unsigned char* src = A.W.tcblob + (size_t)tower * TC_BLOB_BYTES;
        const uint32_t ring = smem_u32(smem + OFF_RING);
        for (int off = 0; off < TC_BLOB_BYTES; off += 6 * B_KSTEP) {          // whole blob HBM -> L2 up front: the ring then runs at L2-hit latency
            const uint32_t bytes = min(6 * B_KSTEP, TC_BLOB_BYTES - off);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + off), "r"(bytes) : "memory");
        }
        for (int i = 0; i < KS1 + KS2 + KSH; ++i) {
            const int s = i % NST, use = i / NST;
            if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
            const uint32_t bytes = i < KS1 + KS2 ? B_KSTEP : H_KSTEP;
            mbar_expect_tx(&full[s], bytes);
            bulk_g2s(ring + s * B_KSTEP, src, bytes, &full[s]);
            src += bytes;
            if (i == NST - 1) TC_MARK(14);
        }
        TC_MARK(15);
    } else if (t == NWORK + 32) {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t a1 = smem_u32(sA1), a2 = smem_u32(sA2), ring = smem_u32(smem + OFF_RING);
        int i = 0;
        for (int layer = 0; layer < 3; ++layer) {
            mbar_wait(a_ready, layer & 1);
            tc_fence_after();
            TC_MARK(layer == 0 ? 2 : layer == 1 ? 7 : 11);
            const int nks = layer == 0 ? KS1 : layer == 1 ? KS2 : KSH;
            const uint32_t d = tmem + layer * NG;
            const uint32_t idesc = layer == 2 ? make_idesc(TM, NHEAD) : make_idesc(TM, NG);
            const uint32_t bhalf = layer == 2 ? H_HALF : B_HALF, blbo = layer == 2 ? NHEAD * 16 : NG * 16;
            for (int ks = 0; ks < nks; ++ks, ++i) {
                const int s = i % NST;
                mbar_wait(&full[s], (i / NST) & 1);
                tc_fence_after();
                if (i == 0) TC_MARK(3);
                const uint32_t ab = (layer == 1 && ks >= 6) ? a2 + (ks - 6) * A_KSTEP : a1 + ks * A_KSTEP;
                const uint32_t bb = ring + s * B_KSTEP;
                const uint64_t ah = make_desc(ab, TM * 16, 128), al = make_desc(ab + A_HALF, TM * 16, 128);
                const uint64_t bh = make_desc(bb, blbo, 128), bl = make_desc(bb + bhalf, blbo, 128);
                mma_tf32(d, al, bh, idesc, ks > 0);
                mma_tf32(d, ah, bl, idesc, 1);
                mma_tf32(d, ah, bh, idesc, 1);
                umma_commit(&empty[s]);
            }
            umma_commit(&d_full[layer]);
            TC_MARK(layer == 0 ? 4 : layer == 1 ? 8 : 13);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWORK / 32 + 1) { __syncwarp(); tmem_dealloc(tmem, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------------------
// Bring-up / regression probe: D[128 x N] = A[128 x K] * B[N x K]^T through the same descriptors, split and TMEM path.
// variant bit 0: single tf32 pass (no split) -- the accuracy gap to the split result proves the three-product path is live.
__global__ void __launch_bounds__(128, 1) tc_gemm_probe_kernel(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ Dg, int K, int N, int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int nks = K / 8;
    unsigned char* sA = smem;
    unsigned char* sB = smem + nks * A_KSTEP;
    const int bkstep = 4 * N * 16;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + nks * bkstep);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(tmem_ptr, 256); tmem_relinquish(); }
    for (int idx = t; idx < TM * (K / 4); idx += 128) {
        const int r = idx % TM, q = idx / TM;
        const float v[4] = {Ag[r * K + 4 * q], Ag[r * K + 4 * q + 1], Ag[r * K + 4 * q + 2], Ag[r * K + 4 * q + 3]};
        store_hilo(sA, op_off(4 * q, r, TM), A_HALF, v);
    }
    for (int idx = t; idx < N * (K / 4); idx += 128) {
        const int r = idx % N, q = idx / N;
        const float v[4] = {Bg[r * K + 4 * q], Bg[r * K + 4 * q + 1], Bg[r * K + 4 * q + 2], Bg[r * K + 4 * q + 3]};
        store_hilo(sB, op_off(4 * q, r, N), 2 * N * 16, v);
    }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (t == 0) {
        const uint32_t idesc = make_idesc(TM, N);
        for (int ks = 0; ks < nks; ++ks) {
            const uint32_t ab = smem_u32(sA) + ks * A_KSTEP, bb = smem_u32(sB) + ks * bkstep;
            const uint32_t albo = TM * 16, asbo = 128, blbo = N * 16, bsbo = 128;
            const uint64_t ah = make_desc(ab, albo, asbo), al = make_desc(ab + A_HALF, albo, asbo);
            const uint64_t bh = make_desc(bb, blbo, bsbo), bl = make_desc(bb + 2 * N * 16, blbo, bsbo);
            if (variant & 1) {
                mma_tf32(tmem, ah, bh, idesc, ks > 0);
            } else {
                mma_tf32(tmem, al, bh, idesc, ks > 0);
                mma_tf32(tmem, ah, bl, idesc, 1);
                mma_tf32(tmem, ah, bh, idesc, 1);
            }
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) Dg[(size_t)(32 * warp + lane) * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 256); }
}

}  // namespace tc

void launch_lstm_act_tc(const ActArgs& a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(tc::lstm_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES); configured = true; }
    dim3 grid((a.N + tc::TM - 1) / tc::TM, 2);
    tc::lstm_act_tc_kernel<<<grid, tc::NTHR, tc::SMEM_BYTES, st>>>(a);
}

void tc_timeline(int enable, long long* out16) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, tc::g_timeline, sizeof(long long) * 16);
    cudaMemcpyToSymbol(tc::g_timeline_on, &enable, sizeof(int));
}

int launch_tc_gemm_probe(const float* dA, const float* dB, float* dD, int K, int N, int variant, cudaStream_t st) {
    if (K % 8 || K < 8 || K > 64 || N % 16 || N < 16 || N > 256) return -1;
    const int bytes = (K / 8) * (tc::A_KSTEP + 4 * N * 16) + 64;
    if (cudaFuncSetAttribute(tc::tc_gemm_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return -2;
    tc::tc_gemm_probe_kernel<<<1, 128, bytes, st>>>(dA, dB, dD, K, N, variant);
    return 0;
}

// Host-side packing of one tower's weights into the streamed layout (called by irrl_policy_set_params):
// k-steps of layer 0 (11), layer 1 (12) and the head (6); each = [hi: chunk(2) x row x 16 B][lo: same].
// Row n of a gate block is column 4*unit + gate of the reference's [in,192] matrices (their column = gate*48 + unit).
void pack_tc_blob(const float* wx0, const float* wh0, const float* wx1, const float* wh1, const float* head_w, int head_cols, unsigned char* out) {
    auto put = [](unsigned char* base, uint32_t off, uint32_t lo_off, float v) {
        float hi = tc::tf32_hi(v); float lo = v - hi;
        memcpy(base + off, &hi, 4); memcpy(base + off + lo_off, &lo, 4);
    };
    memset(out, 0, TC_BLOB_BYTES);
    for (int n = 0; n < tc::NG; ++n) {
        const int src = (n & 3) * LSTM_H + (n >> 2);
        for (int k = 0; k < OB_DIM; ++k) put(out, tc::op_off(k, n, tc::NG), tc::B_HALF, wx0[(size_t)k * tc::NG + src]);
        for (int k = 0; k < LSTM_H; ++k) put(out, tc::op_off(40 + k, n, tc::NG), tc::B_HALF, wh0[(size_t)k * tc::NG + src]);
        unsigned char* l1 = out + tc::KS1 * tc::B_KSTEP;
        for (int k = 0; k < LSTM_H; ++k) put(l1, tc::op_off(k, n, tc::NG), tc::B_HALF, wx1[(size_t)k * tc::NG + src]);
        for (int k = 0; k < LSTM_H; ++k) put(l1, tc::op_off(48 + k, n, tc::NG), tc::B_HALF, wh1[(size_t)k * tc::NG + src]);
    }
    unsigned char* hd = out + (tc::KS1 + tc::KS2) * tc::B_KSTEP;
    for (int n = 0; n < head_cols; ++n)
        for (int k = 0; k < LSTM_H; ++k) put(hd, tc::op_off(k, n, tc::NHEAD), tc::H_HALF, head_w[(size_t)k * head_cols + n]);
}

}  // namespace irrl
