// Tensor-core (tcgen05 + TMEM + bulk-copy) version of the fused two-tower 2x48 LSTM policy act() for sm_100a.
//
// Same contract as lstm_act_kernel (policy_kernels.cu; CustomLSTMPolicy.step run_bp_v5.py:178-185 over the graph
// run_bp_v5.py:143-176): both layers of one tower, the pi / V head, the Gaussian sample and neglogp in one launch.
//
// Mapping.  A CTA owns 128 environments (the MMA M dimension, one TMEM lane per environment) and one tower
// (blockIdx.y).  Every layer is one GEMM  gates[128 x 192] = [x ; h][128 x K] * W^T[K x 192]  issued by one elected lane
// as tcgen05.mma.kind::tf32 instructions (M128 N192 K8) with the fp32 accumulator in tensor memory.
//   * fp32 accuracy on tf32 hardware: both operands are split x = hi + lo (hi = nearest tf32 number, lo = x - hi, both
//     exact in fp32) and three products are accumulated, lo*hi + hi*lo + hi*hi ("3xTF32"); the dropped lo*lo term is
//     2^-24 relative.  Agreement with the fp32 FMA kernel and the float64 oracle stays inside the 2e-5 bar
//     (tests/test_gpu_policy_parity.py).
//   * operands live in shared memory in the UMMA canonical K-major no-swizzle layout
//     [k-step of 8][hi|lo][16-byte k-chunk (2)][row][4 floats]  (core matrix = 8 rows x 16 B, SBO = 128 B,
//     LBO = rows * 16 B), so a thread that owns one environment row writes its 4 new hidden units as one float4.
//     The measured MMA rate does not depend on the layout type (irrl_tc_mma_rate: 98 cycles per M128 N192 K8).
//   * the weights are stored once (irrl_policy_set_params) in exactly that layout, hi and lo pre-split, and stream
//     L2 -> shared memory as 12 KB cp.async.bulk copies completing on mbarriers (TMA engine) through a ring of 3 slots,
//     6 for layer 1 (3 more slots overlay the part of the layer-0 A tile that is dead by then); tcgen05.commit
//     releases a slot as soon as its three MMAs have read it.  Biases / logstd ride in the same blob.
//   * recurrent state: one 384-byte bulk copy per environment row and layer ([c | h], contiguous in the [N,384] state)
//     into a padded staging tile (pitch 400 B: conflict-free for row-per-lane access), and one bulk store per row back.
//     No thread ever issues a strided global access; global stores are only issued behind a hand-off because
//     fence.proxy.async is a full memory barrier for the issuing thread.
//   * epilogue: 16 warps; warp w reads TMEM lanes 32 (w%4).., gate columns 48 (w/4).. with tcgen05.ld.32x32b.x16 --
//     columns are gate-interleaved (4 unit + gate), so 16 columns = 4 hidden units x (i,f,o,g): the cell update is done
//     in registers (5 ex2 + 3 rcp per unit: MUFU-bound), h (hi/lo) goes straight into the A-operand tile of the next GEMM.
//   * the heads (48 x 12 means, 48 x 1 value) run on the CUDA cores: every epilogue thread multiplies its 12 top-layer units with the
//     fp32 head weights (one 3 KB copy over the dead layer-1 tile), the four column blocks of a row meet in shared memory; the
//     Gaussian draws, sigma and neglogp are computed beforehand; lanes of warps 0-3 then write their own environment.
//     (A third tcgen05 GEMM with N = 16 was measured first: 4.7 k cycles of hand-offs for 18 tiny MMAs.)
// Roles: warps 0-15 stage + epilogue, warp 16 lane 0 = bulk-copy producer, warp 17 = MMA issuer (owns TMEM).
// Measured (B200, L2 flushed): 25 us at 4096 and 8192 environments, 44 us at 16384, 80 us at 32768
// (fp32 FMA kernel: 40 / 57 / 103 / 174 us).
#include <cstring>
#include "env_device.cuh"
#include "env_kernels.h"
#include "tc_common.cuh"

namespace irrl {
namespace tc {

constexpr int TM = 128;                          // environments per CTA = MMA M
constexpr int NG = 192;                          // gate columns = MMA N
constexpr int KS1 = 11, KS2 = 12;                // k-steps (8 floats) of layer 0 [obs 35 -> 40 ; h 48] and layer 1 [h_new 48 ; h 48]
constexpr int A_HALF = 2 * TM * 16;              // bytes of the hi (or lo) part of one A k-step : [chunk 2][row 128][16 B]
constexpr int A_KSTEP = 2 * A_HALF;              // 8192
constexpr int B_HALF = 2 * NG * 16;              // 6144
constexpr int B_KSTEP = 2 * B_HALF;              // 12288
constexpr int NST = 3;                           // weight ring stages in their own shared memory
constexpr int NSLOT = 6;                         // + 3 more for layer 1, carved out of the part of the layer-0 A tile that is dead by then
constexpr int NWORK = 512;                       // staging / epilogue threads (16 warps: TMEM lane quarter w%4, 48-column block w/4)
constexpr int NTHR = NWORK + 64;
constexpr int TMEM_COLS = 512;                   // 192 (layer 0) + 192 (layer 1) -> next power of two
constexpr int STG_PITCH = 400;                   // bytes per row of the state staging tile: [c(48) h(48)] = 384 B + 16 B (odd number of 16-byte groups: no bank conflicts)
constexpr int OFF_A1 = 0;
constexpr int OFF_A2 = OFF_A1 + KS1 * A_KSTEP;   // 90112 : h(t-1) of layer 1 (and, before that, the raw observation tile)
constexpr int OFF_RING = OFF_A2 + 6 * A_KSTEP;   // 139264
constexpr int OFF_STG = OFF_RING + NST * B_KSTEP;    // 176128 : [row][c | h] of one layer, in (bulk loads) and out (bulk stores)
constexpr int OFF_BIAS = OFF_STG + TM * STG_PITCH;   // 227328 : TC_BIAS_BYTES = 2 x 192 gate biases, 16 head biases, 16 logstd (+ pad)
constexpr int OFF_BAR = OFF_BIAS + TC_BIAS_BYTES;
constexpr int NBAR = 2 * NSLOT + 1 + 3 + 4;      // full, empty, a_ready, d_full[3], in_bar[2], obs_bar, bias_bar
constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
static_assert((KS1 + KS2) * B_KSTEP + TC_BIAS_BYTES == TC_BLOB_BYTES, "blob size");
constexpr int BIAS_HEADW = 2 * NG + 48;          // float offset of the [48][12] head weights inside the bias block
constexpr int OFF_HEADP = 0;                     // inside the (dead by then) layer-1 A tile: [4 column blocks][128 rows][12] partial head sums
static_assert(SMEM_BYTES <= 232448, "shared memory");
static_assert(6 * A_KSTEP + 3 * B_KSTEP <= KS1 * A_KSTEP, "extra ring slots must fit behind the h_new columns of the layer-0 A tile");
// weight k-step i (0..22) lives in ring slot: layer 0 cycles through the 3 dedicated slots, layer 1 through all 6
__device__ __forceinline__ int slot_of(int i) { return i < KS1 ? i % NST : (i - KS1) % NSLOT; }
__device__ __forceinline__ uint32_t slot_addr(uint32_t ring, uint32_t a1, int s) { return s < NST ? ring + s * B_KSTEP : a1 + 6 * A_KSTEP + (s - NST) * B_KSTEP; }

// optional timeline of CTA (0,0) in SM clocks (irrl_tc_timeline): a cheap way to see which hand-off paces the kernel
__device__ long long g_timeline[32];
__device__ int g_timeline_on = 0;
#define TC_MARK(slot) do { if (dbg) g_timeline[slot] = clock64(); } while (0)      // dbg is read once per thread

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); }

__global__ void __launch_bounds__(NTHR, 1) lstm_act_tc_kernel(const __grid_constant__ ActArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tower = blockIdx.y, e0 = blockIdx.x * TM;
    const int dbg = (blockIdx.x == 0 && blockIdx.y == 0) ? g_timeline_on : 0;
    if (t == 0) TC_MARK(23);
    unsigned char* sA1 = smem + OFF_A1;
    unsigned char* sA2 = smem + OFF_A2;
    unsigned char* sSTG = smem + OFF_STG;
    const float* sbias = reinterpret_cast<const float*>(smem + OFF_BIAS);      // [2][192] gate biases, [16] head bias, [16] logstd
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty = full + NSLOT;
    uint64_t* a_ready = empty + NSLOT;
    uint64_t* d_full = a_ready + 1;
    uint64_t* in_bar = d_full + 3;          // [2] : state rows of layer 0 / layer 1 have landed in the staging tile
    uint64_t* obs_bar = in_bar + 2;
    uint64_t* bias_bar = obs_bar + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bias_bar + 1);
    const int rows = min(TM, A.N - e0);
    // the observation tile is one contiguous block; the bulk engine wants 16-byte granularity on both ends
    const bool obs_bulk = ((reinterpret_cast<uintptr_t>(A.obs) & 15) == 0) && ((rows & 3) == 0);
    const unsigned char* blob = A.W.tcblob + (size_t)tower * TC_BLOB_BYTES;

    if (t == NWORK) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(a_ready, NWORK);
        for (int i = 0; i < 3; ++i) mbar_init(&d_full[i], 1);
        mbar_init(&in_bar[0], TM); mbar_init(&in_bar[1], TM);
        mbar_init(obs_bar, 1); mbar_init(bias_bar, 1);
        fence_mbar_init();
        // everything that does not depend on another thread is requested right here
        mbar_expect_tx(bias_bar, TC_BIAS_BYTES);
        bulk_g2s(smem_u32(smem + OFF_BIAS), blob + TC_BLOB_BYTES - TC_BIAS_BYTES, TC_BIAS_BYTES, bias_bar);
        if (obs_bulk) {
            mbar_expect_tx(obs_bar, rows * OB_DIM * 4);
            bulk_g2s(smem_u32(sA2), A.obs + (size_t)e0 * OB_DIM, rows * OB_DIM * 4, obs_bar);
        }
    }
    if (warp == NWORK / 32 + 1) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (t == 0) TC_MARK(0);

    if (t < NWORK) {
        const int wq = warp & 3, part = warp >> 2, row = 32 * wq + lane, env = e0 + row, envc = min(env, A.N - 1);
        const bool valid = env < A.N;
        float* strow = A.state + (size_t)min(e0 + t, A.N - 1) * LSTM_STATE + tower * 4 * LSTM_H;     // row t of the tile (t < 128 only)
        // ------------------------------------------------------------ state rows [c0 h0] -> staging tile (one 384-byte bulk copy per row)
        if (t < TM) { mbar_expect_tx(&in_bar[0], 2 * LSTM_H * 4); bulk_g2s(smem_u32(sSTG + t * STG_PITCH), strow, 2 * LSTM_H * 4, &in_bar[0]); }
        const float keep = (A.done && A.done[envc]) ? 0.f : 1.f;
        float kp[3];
#pragma unroll
        for (int it = 0; it < 3; ++it) { const int r = (t + it * NWORK) & (TM - 1); kp[it] = (A.done && A.done[min(e0 + r, A.N - 1)]) ? 0.f : 1.f; }
        float* stage = reinterpret_cast<float*>(sA2);                 // layer-1 tile doubles as the observation staging buffer until h1 is stored
        if (obs_bulk) {
            mbar_wait(obs_bar, 0);
        } else {                                                      // unaligned / ragged tile: plain loads
            for (int i = t; i < TM * OB_DIM; i += NWORK) stage[i] = (i < rows * OB_DIM) ? __ldg(A.obs + (size_t)e0 * OB_DIM + i) : 0.f;
            worker_sync();
        }
        if (t == 0) TC_MARK(16);
#pragma unroll
        for (int it = 0; it < 3; ++it) {                             // observations: item = (row, 4 k), zero padded 35 -> 40, rows past N zero
            const int idx = t + it * NWORK, r = idx & (TM - 1), q = idx >> 7;
            if (q < 10) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (4 * q + j < OB_DIM && r < rows) ? stage[r * OB_DIM + 4 * q + j] : 0.f;
                store_hilo(sA1, op_off(4 * q, r, TM), A_HALF, v);
            }
        }
        if (t == 0) TC_MARK(17);
        mbar_wait(&in_bar[0], 0);
        if (t == 0) TC_MARK(18);
#pragma unroll
        for (int it = 0; it < 3; ++it) {                             // h(t-1) of layer 0, masked (SB lstm(): h *= 1-m)
            const int idx = t + it * NWORK, r = idx & (TM - 1), q = idx >> 7;
            const float4 h = *reinterpret_cast<const float4*>(sSTG + r * STG_PITCH + (LSTM_H + 4 * q) * 4);
            const float v0[4] = {h.x * kp[it], h.y * kp[it], h.z * kp[it], h.w * kp[it]};
            store_hilo(sA1, op_off(40 + 4 * q, r, TM), A_HALF, v0);
        }
        float4 ca, cb, cc;                                            // c(t-1) of this lane's 12 units (TMEM lane = row)
        ca = *reinterpret_cast<const float4*>(sSTG + row * STG_PITCH + (12 * part) * 4);
        cb = *reinterpret_cast<const float4*>(sSTG + row * STG_PITCH + (12 * part + 4) * 4);
        cc = *reinterpret_cast<const float4*>(sSTG + row * STG_PITCH + (12 * part + 8) * 4);
        if (t == 0) TC_MARK(19);
        fence_proxy_async();
        mbar_arrive(a_ready);
        if (t == 0) TC_MARK(1);
        worker_sync();                                                // staging tile and observation buffer are free again
        // ---- behind the first GEMM: rows [c1 h1] -> staging tile -> A tile of layer 1 / registers; rollout copies of obs and mask
        if (t < TM) { mbar_expect_tx(&in_bar[1], 2 * LSTM_H * 4); bulk_g2s(smem_u32(sSTG + t * STG_PITCH), strow + 2 * LSTM_H, 2 * LSTM_H * 4, &in_bar[1]); }
        if (tower == 0 && A.obs_store) {                              // mb_obs (ppo2.py:522): global stores only ever follow a hand-off,
            for (int i = t; i < rows * OB_DIM; i += NWORK) A.obs_store[(size_t)e0 * OB_DIM + i] = stage[i];      // fence.proxy.async would wait for them
        }
        if (tower == 0 && A.done_store && t < rows) A.done_store[e0 + t] = A.done ? A.done[e0 + t] : 0;       // mb_dones (ppo2.py:526)
        // ---- also behind the first GEMM: everything of the Gaussian policy head that does not depend on the network output
        mbar_wait(bias_bar, 0);
        const float* hb = sbias + 2 * NG;
        float sg[3][4];                                               // sigma * z per action: the 12 Gaussian draws, sigma = exp(logstd) and
        float nlp = 0.5f * 1.8378770664093453f * ACT_DIM;             // neglogp = sum(0.5 z^2 + logstd) + 6 ln(2 pi) (z = the draw itself)
        if (part == 0 && tower == 0) {
#pragma unroll
            for (int g3 = 0; g3 < 3; ++g3) {
                float g[4] = {0.f, 0.f, 0.f, 0.f};
                if (!A.deterministic) gauss4(A.seed, (uint32_t)envc + A.env_offset, A.tick, P_POLICY_EPS + g3, g);
#pragma unroll
                for (int j = 0; j < 4; ++j) { sg[g3][j] = hb[32 + 4 * g3 + j] * g[j]; nlp += 0.5f * g[j] * g[j] + hb[16 + 4 * g3 + j]; }
            }
        }

        mbar_wait(&in_bar[1], 0);
        worker_sync();                                                // obs_store has finished reading the buffer that h1 overwrites
        float4 da, db, dc;
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int idx = t + it * NWORK, r = idx & (TM - 1), q = idx >> 7;
            const float4 h = *reinterpret_cast<const float4*>(sSTG + r * STG_PITCH + (LSTM_H + 4 * q) * 4);
            const float v1[4] = {h.x * kp[it], h.y * kp[it], h.z * kp[it], h.w * kp[it]};
            store_hilo(sA2, op_off(4 * q, r, TM), A_HALF, v1);
        }
        da = *reinterpret_cast<const float4*>(sSTG + row * STG_PITCH + (12 * part) * 4);
        db = *reinterpret_cast<const float4*>(sSTG + row * STG_PITCH + (12 * part + 4) * 4);
        dc = *reinterpret_cast<const float4*>(sSTG + row * STG_PITCH + (12 * part + 8) * 4);
        worker_sync();                                                // staging tile free for the epilogue's output rows

        // ------------------------------------------------------------ cell updates straight out of tensor memory
        // Rolled loops on purpose: every warp runs this code once per launch, so its size is paid in instruction fetches.
        float4 ha = make_float4(0.f, 0.f, 0.f, 0.f), hb4 = ha, hc = ha;
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {
            mbar_wait(&d_full[l], 0);
            tc_fence_after();
            if (t == 0) TC_MARK(5 + 4 * l);
            const uint32_t taddr = tmem + ((uint32_t)(32 * wq) << 16) + l * NG + 48 * part;
            const float* bl = sbias + l * NG + 48 * part;
            unsigned char* orow = sSTG + row * STG_PITCH + (12 * part) * 4;
#pragma unroll 1
            for (int i = 0; i < 3; ++i) {
                uint32_t v[16];
                tmem_ld16(taddr + 16 * i, v);
                tmem_ld_wait();
                const float cold[4] = {ca.x, ca.y, ca.z, ca.w};
                float cn[4], hn[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {      // gate order i,f,o,g (CustomerLstmNN.py:119-126); reciprocals shared pairwise
                    const float4 b4 = *reinterpret_cast<const float4*>(bl + 16 * i + 4 * j);
                    const float zi = __uint_as_float(v[4 * j + 0]) + b4.x, zf = __uint_as_float(v[4 * j + 1]) + b4.y;
                    const float zo = __uint_as_float(v[4 * j + 2]) + b4.z, zg = __uint_as_float(v[4 * j + 3]) + b4.w;
                    const float ei = __expf(-zi), ef = __expf(-zf), eo = __expf(-zo), eg = __expf(-2.0f * fabsf(zg));
                    const float ig_gg = copysignf(__fdividef(1.0f - eg, (1.0f + ei) * (1.0f + eg)), zg);       // sigmoid(zi) * tanh(zg)
                    const float fg = __fdividef(1.0f, 1.0f + ef);
                    cn[j] = fmaf(fg, cold[j] * keep, ig_gg);
                    const float ec = __expf(-2.0f * fabsf(cn[j]));
                    hn[j] = copysignf(__fdividef(1.0f - ec, (1.0f + eo) * (1.0f + ec)), cn[j]);                 // sigmoid(zo) * tanh(c)
                }
                if (l == 0) store_hilo(sA1, op_off(12 * part + 4 * i, row, TM), A_HALF, hn);      // A operand of the next GEMM (k = unit index)
                *reinterpret_cast<float4*>(orow + 16 * i) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                *reinterpret_cast<float4*>(orow + LSTM_H * 4 + 16 * i) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                ca = cb; cb = cc;
                ha = hb4; hb4 = hc; hc = make_float4(hn[0], hn[1], hn[2], hn[3]);      // after the loop: units 12 part + 0..3 | 4..7 | 8..11
            }
            tc_fence_before();
            fence_proxy_async();
            if (l == 0) mbar_arrive(a_ready);
            if (t == 0) TC_MARK(6 + 4 * l);
            worker_sync();                                            // all output rows of this layer are in the staging tile
            if (t < rows) { bulk_s2g(strow + l * 2 * LSTM_H, smem_u32(sSTG + t * STG_PITCH), 2 * LSTM_H * 4); bulk_commit(); }
            if (l == 0) {
                ca = da; cb = db; cc = dc;
                if (t < rows) bulk_wait_read0();                      // the engine has read the rows: the tile may be overwritten
                worker_sync();
            }
        }
        // ------------------------------------------------------------ heads on the CUDA cores: 48 x 12 (48 x 1) is too small to pay
        // for another operand hand-off to the tensor core.  Every thread multiplies its 12 hidden units of the top layer with the
        // head weights (fp32 [48][16], copied over the dead layer-1 A tile), the 4 column blocks of a row are summed through shared memory.
        {
            const float4* Wh = reinterpret_cast<const float4*>(sbias + BIAS_HEADW) + (12 * part) * 3;       // row u = 3 float4 (12 outputs)
            const float hu[12] = {ha.x, ha.y, ha.z, ha.w, hb4.x, hb4.y, hb4.z, hb4.w, hc.x, hc.y, hc.z, hc.w};
            float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0, p2 = p0;
#pragma unroll
            for (int u = 0; u < 12; ++u) {
                const float4 w0 = Wh[3 * u], w1 = Wh[3 * u + 1], w2 = Wh[3 * u + 2];
                p0.x = fmaf(hu[u], w0.x, p0.x); p0.y = fmaf(hu[u], w0.y, p0.y); p0.z = fmaf(hu[u], w0.z, p0.z); p0.w = fmaf(hu[u], w0.w, p0.w);
                p1.x = fmaf(hu[u], w1.x, p1.x); p1.y = fmaf(hu[u], w1.y, p1.y); p1.z = fmaf(hu[u], w1.z, p1.z); p1.w = fmaf(hu[u], w1.w, p1.w);
                p2.x = fmaf(hu[u], w2.x, p2.x); p2.y = fmaf(hu[u], w2.y, p2.y); p2.z = fmaf(hu[u], w2.z, p2.z); p2.w = fmaf(hu[u], w2.w, p2.w);
            }
            float4* P = reinterpret_cast<float4*>(sA2 + OFF_HEADP) + (size_t)(part * TM + row) * 3;
            P[0] = p0; P[1] = p1; P[2] = p2;
        }
        worker_sync();
        if (t == 0) TC_MARK(12);
        if (part == 0) {
            float v[12];
            {
                const float4* P = reinterpret_cast<const float4*>(sA2 + OFF_HEADP) + (size_t)row * 3;
                float4 s0 = P[0], s1 = P[1], s2 = P[2];
#pragma unroll
                for (int pb = 1; pb < 4; ++pb) {
                    const float4 q0 = P[pb * TM * 3], q1 = P[pb * TM * 3 + 1], q2 = P[pb * TM * 3 + 2];
                    s0.x += q0.x; s0.y += q0.y; s0.z += q0.z; s0.w += q0.w; s1.x += q1.x; s1.y += q1.y; s1.z += q1.z; s1.w += q1.w;
                    s2.x += q2.x; s2.y += q2.y; s2.z += q2.z; s2.w += q2.w;
                }
                v[0] = s0.x; v[1] = s0.y; v[2] = s0.z; v[3] = s0.w; v[4] = s1.x; v[5] = s1.y; v[6] = s1.z; v[7] = s1.w; v[8] = s2.x; v[9] = s2.y; v[10] = s2.z; v[11] = s2.w;
            }
            if (tower == 0) {
                if (valid) {
#pragma unroll
                    for (int g3 = 0; g3 < 3; ++g3) {
                        float x[4], mean[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { mean[j] = v[4 * g3 + j] + hb[4 * g3 + j]; x[j] = mean[j] + sg[g3][j]; }
                        reinterpret_cast<float4*>(A.action + (size_t)env * ACT_DIM)[g3] = make_float4(x[0], x[1], x[2], x[3]);
                        if (A.clipped)                                                                                   // ppo2.py:529-531
                            reinterpret_cast<float4*>(A.clipped + (size_t)env * ACT_DIM)[g3] =
                                make_float4(fminf(fmaxf(x[0], -1.f), 1.f), fminf(fmaxf(x[1], -1.f), 1.f), fminf(fmaxf(x[2], -1.f), 1.f), fminf(fmaxf(x[3], -1.f), 1.f));
                        if (A.mean) reinterpret_cast<float4*>(A.mean + (size_t)env * ACT_DIM)[g3] = make_float4(mean[0], mean[1], mean[2], mean[3]);
                    }
                    A.neglogp[env] = nlp;
                }
            } else if (valid) {
                A.value[env] = v[0] + hb[0];
            }
        }
        if (t == 0) TC_MARK(21);
        if (t < rows) bulk_wait0();                                   // state rows are in HBM before the CTA retires
        if (t == 0) TC_MARK(22);
    } else if (t == NWORK) {
        // ------------------------------------------------------------ weight producer: bulk copies through the ring
        const unsigned char* src = blob;
        const uint32_t ring = smem_u32(smem + OFF_RING), a1 = smem_u32(sA1);
        uint32_t used = 0, phase = 0;                                 // per-slot: has been filled before / parity of its next `empty` completion
        for (int i = 0; i < KS1 + KS2; ++i) {
            const int sl = slot_of(i);
            if (i == KS1 + NST) mbar_wait(&d_full[0], 0);             // slots 3-5 overlay A-tile columns the first GEMM reads
            if (used & (1u << sl)) { mbar_wait(&empty[sl], (phase >> sl) & 1); phase ^= 1u << sl; }
            used |= 1u << sl;
            mbar_expect_tx(&full[sl], B_KSTEP);
            bulk_g2s(slot_addr(ring, a1, sl), src, B_KSTEP, &full[sl]);
            src += B_KSTEP;
            if (i == NST - 1) TC_MARK(14);
        }
        TC_MARK(15);
    } else if (warp == NWORK / 32 + 1) {
        // ------------------------------------------------------------ MMA issuer: the whole warp walks the loop (warp-uniform control flow keeps
        // descriptors in uniform registers), one elected lane issues
        const uint32_t a1 = smem_u32(sA1), a2 = smem_u32(sA2), ring = smem_u32(smem + OFF_RING);
        constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);        // SBO = 128 B, descriptor version 1
        constexpr uint32_t A_LBO = (uint32_t)(TM * 16 >> 4) << 16, G_LBO = (uint32_t)(NG * 16 >> 4) << 16;
        uint32_t phase = 0;                                           // per-slot parity of the next `full` completion
        int i = 0;
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
            mbar_wait(a_ready, layer & 1);
            tc_fence_after();
            if (lane == 0) TC_MARK(layer == 0 ? 2 : 7);
            const int nks = layer == 0 ? KS1 : KS2;
            const uint32_t d = tmem + layer * NG;
            const uint32_t idesc = make_idesc(TM, NG);
#pragma unroll 1
            for (int ks = 0; ks < nks; ++ks, ++i) {
                const int sl = slot_of(i);
                mbar_wait(&full[sl], (phase >> sl) & 1); phase ^= 1u << sl;
                tc_fence_after();
                const uint32_t bb = slot_addr(ring, a1, sl);
                const uint32_t ab = (layer == 1 && ks >= 6) ? a2 + (ks - 6) * A_KSTEP : a1 + ks * A_KSTEP;
                const uint32_t alo = (ab >> 4) | A_LBO, blo = (bb >> 4) | G_LBO;
                const uint64_t ah = ((uint64_t)DESC_HI << 32) | alo, al = ((uint64_t)DESC_HI << 32) | (alo + (A_HALF >> 4));
                const uint64_t bh = ((uint64_t)DESC_HI << 32) | blo, bl = ((uint64_t)DESC_HI << 32) | (blo + (B_HALF >> 4));
                if (elect_one()) {
                    mma_tf32(d, al, bh, idesc, ks > 0);
                    mma_tf32(d, ah, bl, idesc, 1);
                    mma_tf32(d, ah, bh, idesc, 1);
                    umma_commit(&empty[sl]);
                    if (ks == nks - 1) umma_commit(&d_full[layer]);
                }
                __syncwarp();
            }
            if (lane == 0) TC_MARK(layer == 0 ? 4 : 8);
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

// Issue-rate probe: `reps` back-to-back M128 x N x K8 tf32 MMAs on one accumulator, operands described with the given
// shared-memory layout type / strides (contents are irrelevant); returns SM cycles from first issue to completion.
__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(int N, int reps, uint32_t layout_type, uint32_t lbo, uint32_t sbo, uint32_t kadv, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar; __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint32_t idesc = make_idesc(TM, N);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
        const uint32_t hi = (sbo >> 4) | (1u << 14) | (layout_type << 29);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t off = (uint32_t)(r & 3) * kadv;
            const uint64_t ad = ((uint64_t)hi << 32) | (((a0 + off) >> 4) | ((lbo >> 4) << 16));
            const uint64_t bd = ((uint64_t)hi << 32) | (((b0 + off) >> 4) | ((lbo >> 4) << 16));
            if (elect_one()) mma_tf32(tmem, ad, bd, idesc, r > 0);
            __syncwarp();
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (t == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 256); }
}

}  // namespace tc

void launch_lstm_act_tc(const ActArgs& a, cudaStream_t st) {
    {   // the attribute is per device (a process may hold envs / policies on several GPUs through the C ABI)
        static unsigned long long configured_devices = 0ull; int dev = 0; cudaGetDevice(&dev);
        if (dev >= 64 || !((configured_devices >> dev) & 1ull)) { cudaFuncSetAttribute(tc::lstm_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES); if (dev < 64) configured_devices |= 1ull << dev; }
    }
    dim3 grid((a.N + tc::TM - 1) / tc::TM, 2);
    tc::lstm_act_tc_kernel<<<grid, tc::NTHR, tc::SMEM_BYTES, st>>>(a);
}

void tc_timeline(int enable, long long* out16) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, tc::g_timeline, sizeof(long long) * 32);
    cudaMemcpyToSymbol(tc::g_timeline_on, &enable, sizeof(int));
}

int launch_tc_mma_rate(int N, int reps, unsigned layout_type, unsigned lbo, unsigned sbo, unsigned kadv, long long* d_out, cudaStream_t st) {
    if (cudaFuncSetAttribute(tc::tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess) return -2;
    tc::tc_mma_rate_kernel<<<1, 128, 64 * 1024, st>>>(N, reps, layout_type, lbo, sbo, kadv, d_out);
    return 0;
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
void pack_tc_blob(const float* wx0, const float* wh0, const float* b0, const float* wx1, const float* wh1, const float* b1, const float* head_w, const float* head_b,
                  int head_cols, const float* logstd, unsigned char* out) {
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
    float* bias = reinterpret_cast<float*>(out + TC_BLOB_BYTES - TC_BIAS_BYTES);      // [2][192] gate biases, [16] head bias, [16] logstd
    for (int n = 0; n < tc::NG; ++n) { const int src = (n & 3) * LSTM_H + (n >> 2); bias[n] = b0[src]; bias[tc::NG + n] = b1[src]; }
    for (int n = 0; n < head_cols; ++n) bias[2 * tc::NG + n] = head_b[n];
    if (logstd) for (int n = 0; n < ACT_DIM; ++n) { bias[2 * tc::NG + 16 + n] = logstd[n]; bias[2 * tc::NG + 32 + n] = expf(logstd[n]); }
    for (int n = 0; n < head_cols; ++n)                                            // fp32 [48 units][12 outputs], zero padded
        for (int k = 0; k < LSTM_H; ++k) bias[tc::BIAS_HEADW + k * 12 + n] = head_w[(size_t)k * head_cols + n];
}

}  // namespace irrl
