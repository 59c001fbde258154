// Fused two-tower 2x48 LSTM policy act() and GAE for sm_100a.
//
// Replaces CustomLSTMPolicy.step (run_bp_v5.py:178-185 over the graph run_bp_v5.py:143-176: stable-baselines
// `lstm`, `linear`, DiagGaussian sample + neglogp) and Runner's GAE loop (ppo2.py:554-568).
// One launch = both towers x both layers + pi/V heads + Gaussian sample + neglogp + done-mask state reset.
//
// Mapping (fp32 FMA pipe, agrees with the numpy oracle to ~1e-6): a CTA owns 32 environments and both towers
// (192 threads each).  Every layer is one register-tiled GEMM gates[32 x 192] = [x ; h]^T[32 x 96] W[96 x 192]:
// a thread owns 4 envs x 8 columns, the 8 columns being the i,f,o,g gates of two hidden units (weights are stored
// gate-interleaved once at load time), so the cell update happens in registers with no gate round trip.  Weight
// rows stream L2 -> shared memory in 16-row cp.async stages (double buffered); activations sit transposed and
// XOR-swizzled in shared memory.  The rollout stores of Runner.run (mb_obs, mb_dones) ride along.
#include "env_device.cuh"
#include "env_kernels.h"

namespace irrl {

constexpr int H = LSTM_H;       // 48
constexpr int G4 = 4 * H;       // 192 gate columns per layer
constexpr int TM = 32;          // environments per CTA
constexpr int KP = 96;          // padded reduction length of every layer ([x ; h] rows: 35+48 -> 96, 48+48 = 96)
constexpr int KC = 16;          // rows of W per cp.async stage
constexpr int NTHR = 2 * G4;    // 384 threads: tower = t / 192, then (col-group 0..23) x (env-group 0..7)

// ex2.approx-based forms: relative error ~1e-6, far inside the 2e-5 parity bar (tests/test_gpu_policy_parity.py)
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { float e = __expf(-2.0f * fabsf(x)); float r = __fdividef(1.0f - e, 1.0f + e); return copysignf(r, x); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// env column of the transposed activation tiles is XOR-swizzled by the row so that the transposing stores
// (lanes along k, one env) hit 8 banks instead of 1 while the GEMM's float4 reads stay 16-byte aligned
__device__ __forceinline__ int swz(int k, int env) { return (((env >> 2) ^ (k & 7)) << 2) | (env & 3); }

struct __align__(16) ActSmem {
    float X[2][2][KP][TM];        // [tower][layer] GEMM input, transposed: rows = [x ; h(t-1) masked ; 0-pad]
    float W[2][2][KC][G4];        // [stage][tower] weight rows of the current K chunk (gate-interleaved columns)
    float Hout[2][H][TM];         // [tower] output of the top layer
    float nlp[TM][ACT_DIM];
};

// one LSTM layer for both towers: gates = X^T W (+b) by register tiling (4 envs x 8 columns per thread, the 8 columns
// being the i,f,o,g gates of two hidden units), then the cell update in registers.
__device__ __forceinline__ void lstm_layer(const ActArgs& A, ActSmem& s, int layer, int t, int e0) {
    const int tower = t / G4, tt = t % G4, eg = tt & 7, cg = tt >> 3;
    const float* Wg = A.W.wcat[tower * 2 + layer];       // [KP][192], columns permuted to (unit pair, unit, gate)
    float acc[4][8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(A.W.bperm[tower * 2 + layer] + 8 * cg));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(A.W.bperm[tower * 2 + layer] + 8 * cg + 4));
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc[e][0] = b0.x; acc[e][1] = b0.y; acc[e][2] = b0.z; acc[e][3] = b0.w; acc[e][4] = b1.x; acc[e][5] = b1.y; acc[e][6] = b1.z; acc[e][7] = b1.w; }
    }
    // stage loader: KC x 192 floats per tower = 768 float4 -> 4 per thread of the tower (slot j = tt + 192 i)
    auto load_chunk = [&](int c, int stage) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = tt + i * G4, row = j / (G4 / 4), c4 = j % (G4 / 4);
            cp_async16(&s.W[stage][tower][row][4 * c4], Wg + (size_t)(c * KC + row) * G4 + 4 * c4);
        }
    };
    constexpr int NCH = KP / KC;
    load_chunk(0, 0); cp_async_commit();
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) { load_chunk(c + 1, (c + 1) & 1); cp_async_commit(); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float (*Wc)[G4] = s.W[c & 1][tower];
        const float (*Xc)[TM] = s.X[tower][layer];
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const int k = c * KC + kk;
            const float4 x = *reinterpret_cast<const float4*>(&Xc[k][((eg ^ (k & 7)) << 2)]);
            const float4 w0 = *reinterpret_cast<const float4*>(&Wc[kk][8 * cg]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Wc[kk][8 * cg + 4]);
            const float xe[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[e][0] = fmaf(xe[e], w0.x, acc[e][0]); acc[e][1] = fmaf(xe[e], w0.y, acc[e][1]);
                acc[e][2] = fmaf(xe[e], w0.z, acc[e][2]); acc[e][3] = fmaf(xe[e], w0.w, acc[e][3]);
                acc[e][4] = fmaf(xe[e], w1.x, acc[e][4]); acc[e][5] = fmaf(xe[e], w1.y, acc[e][5]);
                acc[e][6] = fmaf(xe[e], w1.z, acc[e][6]); acc[e][7] = fmaf(xe[e], w1.w, acc[e][7]);
            }
        }
        __syncthreads();
    }
    // cell update (gate order i,f,o,g  CustomerLstmNN.py:119-126); c(t-1) straight from HBM with the done mask
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int el = 4 * eg + e, env = e0 + el, envc = min(env, A.N - 1);
        const float keep = (A.done && A.done[envc]) ? 0.f : 1.f;
        float* st = A.state + (size_t)envc * LSTM_STATE + tower * 4 * H + layer * 2 * H;
        const float2 cold = *reinterpret_cast<const float2*>(st + 2 * cg);
        float cn[2], hn[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float ig = sigmoidf_(acc[e][4 * u + 0]), fg = sigmoidf_(acc[e][4 * u + 1]), og = sigmoidf_(acc[e][4 * u + 2]), gg = tanhf_(acc[e][4 * u + 3]);
            cn[u] = fg * ((u ? cold.y : cold.x) * keep) + ig * gg;
            hn[u] = og * tanhf_(cn[u]);
            const int unit = 2 * cg + u;
            if (layer == 0) s.X[tower][1][unit][swz(unit, el)] = hn[u]; else s.Hout[tower][unit][swz(unit, el)] = hn[u];
        }
        if (env < A.N) {
            *reinterpret_cast<float2*>(st + 2 * cg) = make_float2(cn[0], cn[1]);
            *reinterpret_cast<float2*>(st + H + 2 * cg) = make_float2(hn[0], hn[1]);
        }
    }
}

__global__ void __launch_bounds__(NTHR, 2) lstm_act_kernel(const __grid_constant__ ActArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ActSmem& s = *reinterpret_cast<ActSmem*>(smem_raw);
    const int t = threadIdx.x;
    const int e0 = blockIdx.x * TM;
    const PolicyWeights& W = A.W;
    // ---- stage the inputs transposed: layer-0 rows = [obs(35) ; h0 ; 0], layer-1 rows 48.. = h1 (rows 0..47 come from layer 0)
    for (int i = t; i < TM * OB_DIM; i += NTHR) {          // the tile's observations are one contiguous block
        int e = i / OB_DIM, k = i % OB_DIM; int env = min(e0 + e, A.N - 1);
        float v = A.obs[(size_t)env * OB_DIM + k];
        s.X[0][0][k][swz(k, e)] = v; s.X[1][0][k][swz(k, e)] = v;
        if (A.obs_store && e0 + e < A.N) A.obs_store[(size_t)env * OB_DIM + k] = v;     // mb_obs (ppo2.py:522)
    }
    {   // h(t-1) of the 4 (tower, layer) cells, masked (SB lstm(): h *= 1-m); 16 elements per thread, loads batched
        float hv[16];
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            int i = t + it * NTHR, e = i / (4 * H), r = i % (4 * H), cell = r / H, u = r % H;
            int env = min(e0 + e, A.N - 1);
            hv[it] = A.state[(size_t)env * LSTM_STATE + (cell >> 1) * 4 * H + (cell & 1) * 2 * H + H + u];
        }
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            int i = t + it * NTHR, e = i / (4 * H), r = i % (4 * H), cell = r / H, u = r % H, tower = cell >> 1, layer = cell & 1;
            int env = min(e0 + e, A.N - 1);
            float keep = (A.done && A.done[env]) ? 0.f : 1.f;
            int row = (layer == 0 ? OB_DIM : H) + u;
            s.X[tower][layer][row][swz(row, e)] = hv[it] * keep;
        }
    }
    for (int i = t; i < 2 * (KP - OB_DIM - H) * TM; i += NTHR) {   // zero padding rows 83..95 of layer 0
        int tower = i / ((KP - OB_DIM - H) * TM), r = i % ((KP - OB_DIM - H) * TM);
        s.X[tower][0][OB_DIM + H + r / TM][r % TM] = 0.f;
    }
    if (A.done_store && t < TM && e0 + t < A.N) A.done_store[e0 + t] = A.done ? A.done[e0 + t] : 0;     // mb_dones (ppo2.py:526)
    __syncthreads();
    lstm_layer(A, s, 0, t, e0);
    __syncthreads();
    lstm_layer(A, s, 1, t, e0);
    __syncthreads();
    // ---- heads: pi 48 -> 12 (one (env, action) per thread), V 48 -> 1; Gaussian sample + neglogp
    {
        const int e = t % TM, a = t / TM, env = e0 + e;     // 32 x 12 = 384 threads
        float m = W.pi_b[a];
#pragma unroll 8
        for (int k = 0; k < H; ++k) m = fmaf(s.Hout[0][k][swz(k, e)], __ldg(W.pi_w + k * ACT_DIM + a), m);
        float ls = W.logstd[a], sd = expf(ls), eps = 0.f;
        if (!A.deterministic) {
            float g[4]; gauss4(A.seed, (uint32_t)min(env, A.N - 1) + A.env_offset, A.tick, P_POLICY_EPS + (a >> 2), g);
            eps = g[a & 3];
        }
        float act = fmaf(sd, eps, m);
        float z = (act - m) / sd;
        s.nlp[e][a] = 0.5f * z * z + ls;
        if (env < A.N) {
            A.action[(size_t)env * ACT_DIM + a] = act;
            if (A.clipped) A.clipped[(size_t)env * ACT_DIM + a] = fminf(fmaxf(act, -1.f), 1.f);   // ppo2.py:529-531
            if (A.mean) A.mean[(size_t)env * ACT_DIM + a] = m;
        }
    }
    if (t < TM) {
        const int e = t, env = e0 + e;
        float v = W.vf_b[0];
#pragma unroll 8
        for (int k = 0; k < H; ++k) v = fmaf(s.Hout[1][k][swz(k, e)], __ldg(W.vf_w + k), v);
        if (env < A.N) A.value[env] = v;
    }
    __syncthreads();
    if (t < TM && e0 + t < A.N) {
        float acc = 0.5f * 1.8378770664093453f * ACT_DIM;     // 0.5 * ln(2 pi) * 12
#pragma unroll
        for (int a = 0; a < ACT_DIM; ++a) acc += s.nlp[t][a];
        A.neglogp[e0 + t] = acc;
    }
}

int g_act_path = 0;
void launch_lstm_act(const ActArgs& a, cudaStream_t st) {
    // the tensor-core kernel moves state rows with the bulk-copy engine: 16-byte aligned rows required (always true for [N,384] allocations)
    const bool tc = (g_act_path == 2 || (g_act_path == 0 && (a.n_total > a.N ? a.n_total : a.N) >= 256)) && (reinterpret_cast<uintptr_t>(a.state) & 15) == 0;
    if (tc) launch_lstm_act_tc(a, st); else launch_lstm_act_fma(a, st);
}
void launch_lstm_act_fma(const ActArgs& a, cudaStream_t st) {
    {   // the attribute is per device (a process may hold envs / policies on several GPUs through the C ABI)
        static unsigned long long configured_devices = 0ull; int dev = 0; cudaGetDevice(&dev);
        if (dev >= 64 || !((configured_devices >> dev) & 1ull)) { cudaFuncSetAttribute(lstm_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ActSmem)); if (dev < 64) configured_devices |= 1ull << dev; }
    }
    int grid = (a.N + TM - 1) / TM;
    lstm_act_kernel<<<grid, NTHR, sizeof(ActSmem), st>>>(a);
}

// ------------------------------------------------------------------ GAE (ppo2.py:554-568): one thread per env, reverse scan
__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                           const float* __restrict__ last_values, const uint8_t* __restrict__ last_dones,
                           float* __restrict__ adv, float* __restrict__ ret, int T, int N, float gamma, float lam) {
    int n = blockIdx.x * blockDim.x + threadIdx.x; if (n >= N) return;
    float last = 0.f, nextv = last_values[n], nonterm = 1.0f - (float)last_dones[n];
    for (int t = T - 1; t >= 0; --t) {
        size_t i = (size_t)t * N + n;
        float v = values[i];
        float delta = rewards[i] + gamma * nextv * nonterm - v;
        last = delta + gamma * lam * nonterm * last;
        adv[i] = last; ret[i] = last + v;
        nextv = v; nonterm = 1.0f - (float)dones[i];
    }
}
void launch_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values, const uint8_t* last_dones,
                float* adv, float* ret, int T, int N, float gamma, float lam, cudaStream_t st) {
    gae_kernel<<<(N + 127) / 128, 128, 0, st>>>(rewards, values, dones, last_values, last_dones, adv, ret, T, N, gamma, lam);
}

// ------------------------------------------------------------------ BPTT building blocks for the PPO learner (ppo2.py:136-197)
// The recurrent part of training is a time loop of [h W_h GEMM (cuBLAS) -> fused cell kernel]; these two kernels are the fused
// element-wise halves (forward and backward) so that one LSTM step costs two launches instead of ~30 element-wise ones.
// rows = towers * envs; z / gates are [rows,192] in the checkpoint's gate order i,f,o,g; everything else [rows,48].
__global__ void lstm_pw_fwd_kernel(int rows, int n_env, const float* __restrict__ z, const float* __restrict__ c_prev_masked,
                                   const float* __restrict__ keep_next, float* __restrict__ gates, float* __restrict__ c_out,
                                   float* __restrict__ h_out, float* __restrict__ hm_next, float* __restrict__ cm_next) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x; if (idx >= rows * H) return;
    int r = idx / H, u = idx % H;
    const float* zr = z + (size_t)r * G4;
    float i = 1.f / (1.f + expf(-zr[u])), f = 1.f / (1.f + expf(-zr[H + u])), o = 1.f / (1.f + expf(-zr[2 * H + u])), g = tanhf(zr[3 * H + u]);
    float c = f * c_prev_masked[idx] + i * g, h = o * tanhf(c);
    float* gr = gates + (size_t)r * G4; gr[u] = i; gr[H + u] = f; gr[2 * H + u] = o; gr[3 * H + u] = g;
    c_out[idx] = c; h_out[idx] = h;
    float k = keep_next ? keep_next[r % n_env] : 1.f;          // SB lstm(): c *= 1-m ; h *= 1-m before the next cell
    hm_next[idx] = h * k; cm_next[idx] = c * k;
}
// dh_out [rows,48] = dL/dh_t from above; carry_h = dz_{t+1} W_h^T (gradient w.r.t. the masked h fed to step t+1), carry_c likewise;
// keep_up = keep_{t+1}.  Outputs dz_t [rows,192] and carry_c for step t-1 (gradient w.r.t. the masked c fed to this step).
__global__ void lstm_pw_bwd_kernel(int rows, int n_env, const float* __restrict__ dh_out, const float* __restrict__ carry_h,
                                   const float* __restrict__ carry_c, const float* __restrict__ keep_up, const float* __restrict__ gates,
                                   const float* __restrict__ c, const float* __restrict__ c_prev_masked, float* __restrict__ dz,
                                   float* __restrict__ dcm_prev) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x; if (idx >= rows * H) return;
    int r = idx / H, u = idx % H;
    float k = keep_up ? keep_up[r % n_env] : 0.f;
    float dh = dh_out[idx] + (carry_h ? carry_h[idx] * k : 0.f);
    float dc = carry_c ? carry_c[idx] * k : 0.f;
    const float* gr = gates + (size_t)r * G4;
    float i = gr[u], f = gr[H + u], o = gr[2 * H + u], g = gr[3 * H + u];
    float tc = tanhf(c[idx]);
    dc += dh * o * (1.f - tc * tc);
    float* dr = dz + (size_t)r * G4;
    dr[u] = dc * g * i * (1.f - i); dr[H + u] = dc * c_prev_masked[idx] * f * (1.f - f);
    dr[2 * H + u] = dh * tc * o * (1.f - o); dr[3 * H + u] = dc * i * (1.f - g * g);
    dcm_prev[idx] = dc * f;
}
// ------------------------------------------------------------------ sequence-persistent BPTT kernels (one launch per layer and direction)
// A CTA owns 32 environments of one tower for ALL T steps: W_h stays in shared memory, h(t-1) in shared memory, c(t-1) in
// registers; per step it adds h W_h to the streamed input projection xw[t] (prefetched one step ahead), applies the cell and
// streams gates / c / h out.  Environments are independent, so there is no inter-CTA dependency and no per-step launch.
// Layouts (time-major, as the rollout stores them): xw, gates, dz [T,K,N,192] in the checkpoint's gate order i,f,o,g;
// C, H, dH [T,K,N,48]; keep [T,N]; c0, h0 [K,N,48]; wh [K,48,192].
// (measured round 2: MUFU-based sigmoid / tanh and a 3-CTA-per-SM register cap changed nothing / cost 30 %: the loops are bound by the
// latency of ~1.5 warps per scheduler, not by the transcendental instructions; libdevice expf / tanhf stay)
constexpr int SEQ_TM = 32;
constexpr int SEQ_THR = 192;        // 24 unit pairs x 8 env groups; thread = 4 envs x 2 units (x 4 gates)

// bias [K,192] (may be null) is added to the streamed projection here, and HM [T,K,N,48] (may be null) receives the MASKED hidden state fed
// into every step (what the weight gradient dW_h needs): both used to be separate full passes over 9.4 GB / 2.4 GB tensors in PyTorch.
__global__ void __launch_bounds__(SEQ_THR) lstm_seq_fwd_kernel(int T, int K, int N, const float* __restrict__ xw, const float* __restrict__ wh,
                                                               const float* __restrict__ c0, const float* __restrict__ h0, const float* __restrict__ keep,
                                                               float* __restrict__ gates, float* __restrict__ Cs, float* __restrict__ Hs,
                                                               const float* __restrict__ bias, float* __restrict__ HM) {
    __shared__ __align__(16) float Ws[H][G4];          // [k][unit pair][unit][gate]
    __shared__ __align__(16) float hT[H][SEQ_TM];      // masked h(t-1), transposed
    const int tower = blockIdx.y, e0 = blockIdx.x * SEQ_TM, t_ = threadIdx.x, eg = t_ & 7, cg = t_ >> 3;
    const float* whk = wh + (size_t)tower * H * G4;
    for (int i = t_; i < H * G4; i += SEQ_THR) { int k = i / G4, col = i % G4, g = col / H, u = col % H; Ws[k][(u >> 1) * 8 + (u & 1) * 4 + g] = whk[i]; }
    float c[4][2];
    int envs[4]; bool valid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        envs[e] = min(e0 + 4 * eg + e, N - 1); valid[e] = (e0 + 4 * eg + e) < N;
        const size_t o = ((size_t)tower * N + envs[e]) * H + 2 * cg;
        const float k0 = keep[envs[e]];                                     // keep[0][env]
        c[e][0] = c0[o] * k0; c[e][1] = c0[o + 1] * k0;
        hT[2 * cg][4 * eg + e] = h0[o] * k0; hT[2 * cg + 1][4 * eg + e] = h0[o + 1] * k0;
        if (HM && valid[e]) *reinterpret_cast<float2*>(HM + o) = make_float2(h0[o] * k0, h0[o + 1] * k0);      // row (t = 0, tower, env)
    }
    float bg[4][2];
#pragma unroll
    for (int g = 0; g < 4; ++g) { bg[g][0] = bias ? bias[tower * G4 + g * H + 2 * cg] : 0.f; bg[g][1] = bias ? bias[tower * G4 + g * H + 2 * cg + 1] : 0.f; }
    // prefetch xw[0]
    float2 xin[4][4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int g = 0; g < 4; ++g) xin[e][g] = *reinterpret_cast<const float2*>(xw + (((size_t)0 * K + tower) * N + envs[e]) * G4 + g * H + 2 * cg);
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        float acc[4][8];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int g = 0; g < 4; ++g) { acc[e][g] = xin[e][g].x + bg[g][0]; acc[e][4 + g] = xin[e][g].y + bg[g][1]; }
        if (t + 1 < T) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int g = 0; g < 4; ++g) xin[e][g] = *reinterpret_cast<const float2*>(xw + (((size_t)(t + 1) * K + tower) * N + envs[e]) * G4 + g * H + 2 * cg);
        }
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float4 x = *reinterpret_cast<const float4*>(&hT[k][4 * eg]);
            const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][8 * cg]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][8 * cg + 4]);
            const float xe[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[e][0] = fmaf(xe[e], w0.x, acc[e][0]); acc[e][1] = fmaf(xe[e], w0.y, acc[e][1]); acc[e][2] = fmaf(xe[e], w0.z, acc[e][2]); acc[e][3] = fmaf(xe[e], w0.w, acc[e][3]);
                acc[e][4] = fmaf(xe[e], w1.x, acc[e][4]); acc[e][5] = fmaf(xe[e], w1.y, acc[e][5]); acc[e][6] = fmaf(xe[e], w1.z, acc[e][6]); acc[e][7] = fmaf(xe[e], w1.w, acc[e][7]);
            }
        }
        __syncthreads();                                                    // everybody has read h(t-1)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float kn = (t + 1 < T) ? keep[(size_t)(t + 1) * N + envs[e]] : 1.f;
            float gi[2], gf[2], go[2], gg[2], hn[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                gi[u] = 1.f / (1.f + expf(-acc[e][4 * u + 0])); gf[u] = 1.f / (1.f + expf(-acc[e][4 * u + 1]));
                go[u] = 1.f / (1.f + expf(-acc[e][4 * u + 2])); gg[u] = tanhf(acc[e][4 * u + 3]);
                c[e][u] = gf[u] * c[e][u] + gi[u] * gg[u];
                hn[u] = go[u] * tanhf(c[e][u]);
            }
            if (valid[e]) {
                const size_t row = ((size_t)t * K + tower) * N + envs[e];
                float* gr = gates + row * G4 + 2 * cg;
                *reinterpret_cast<float2*>(gr) = make_float2(gi[0], gi[1]); *reinterpret_cast<float2*>(gr + H) = make_float2(gf[0], gf[1]);
                *reinterpret_cast<float2*>(gr + 2 * H) = make_float2(go[0], go[1]); *reinterpret_cast<float2*>(gr + 3 * H) = make_float2(gg[0], gg[1]);
                *reinterpret_cast<float2*>(Cs + row * H + 2 * cg) = make_float2(c[e][0], c[e][1]);
                *reinterpret_cast<float2*>(Hs + row * H + 2 * cg) = make_float2(hn[0], hn[1]);
            }
            c[e][0] *= kn; c[e][1] *= kn;                                   // SB lstm(): c *= 1-m ; h *= 1-m before the next cell
            hT[2 * cg][4 * eg + e] = hn[0] * kn; hT[2 * cg + 1][4 * eg + e] = hn[1] * kn;
            if (HM && valid[e] && t + 1 < T) *reinterpret_cast<float2*>(HM + (((size_t)(t + 1) * K + tower) * N + envs[e]) * H + 2 * cg) = make_float2(hn[0] * kn, hn[1] * kn);
        }
        __syncthreads();
    }
}

// backward through time: dz[t] from (dH[t] + carry_h keep[t+1], carry_c keep[t+1]); carry_h = dz[t] W_h^T, carry_c = dc f
__global__ void __launch_bounds__(SEQ_THR) lstm_seq_bwd_kernel(int T, int K, int N, const float* __restrict__ dH, const float* __restrict__ wh,
                                                               const float* __restrict__ c0, const float* __restrict__ keep, const float* __restrict__ gates,
                                                               const float* __restrict__ Cs, float* __restrict__ dz, float* __restrict__ db_part) {
    extern __shared__ __align__(16) unsigned char seq_smem[];          // 60 KB: above the static limit, opt-in dynamic
    float (*WT)[H] = reinterpret_cast<float (*)[H]>(seq_smem);                                   // [permuted col][unit k] = wh[k][col]
    float (*dzT)[SEQ_TM] = reinterpret_cast<float (*)[SEQ_TM]>(seq_smem + sizeof(float) * G4 * H); // dz(t), permuted columns, transposed
    const int tower = blockIdx.y, e0 = blockIdx.x * SEQ_TM, t_ = threadIdx.x, eg = t_ & 7, cg = t_ >> 3;
    const float* whk = wh + (size_t)tower * H * G4;
    for (int i = t_; i < H * G4; i += SEQ_THR) { int k = i / G4, col = i % G4, g = col / H, u = col % H; WT[(u >> 1) * 8 + (u & 1) * 4 + g][k] = whk[i]; }
    int envs[4]; bool valid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { envs[e] = min(e0 + 4 * eg + e, N - 1); valid[e] = (e0 + 4 * eg + e) < N; }
    float dbacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // sum of dz over this thread's envs and all t: the bias gradient (db_part, may be null)
    float ch[4][2], cc[4][2];                           // carries (gradient w.r.t. the masked h / c fed to step t+1)
#pragma unroll
    for (int e = 0; e < 4; ++e) { ch[e][0] = ch[e][1] = cc[e][0] = cc[e][1] = 0.f; }
    // per-step operands, loaded one step ahead
    struct In { float2 dh, c, cp, gi, gf, go, gg; float ku, kt; };
    In cur[4], nxt[4];
    auto load_step = [&](int t, In* in) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const size_t row = ((size_t)t * K + tower) * N + envs[e];
            in[e].ku = (t + 1 < T) ? keep[(size_t)(t + 1) * N + envs[e]] : 0.f;
            in[e].kt = keep[(size_t)t * N + envs[e]];
            in[e].dh = *reinterpret_cast<const float2*>(dH + row * H + 2 * cg);
            in[e].c = *reinterpret_cast<const float2*>(Cs + row * H + 2 * cg);
            in[e].cp = (t > 0) ? *reinterpret_cast<const float2*>(Cs + (((size_t)(t - 1) * K + tower) * N + envs[e]) * H + 2 * cg)
                               : *reinterpret_cast<const float2*>(c0 + ((size_t)tower * N + envs[e]) * H + 2 * cg);
            const float* gr = gates + row * G4 + 2 * cg;
            in[e].gi = *reinterpret_cast<const float2*>(gr); in[e].gf = *reinterpret_cast<const float2*>(gr + H);
            in[e].go = *reinterpret_cast<const float2*>(gr + 2 * H); in[e].gg = *reinterpret_cast<const float2*>(gr + 3 * H);
        }
    };
    load_step(T - 1, cur);
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
        if (t > 0) load_step(t - 1, nxt);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const size_t row = ((size_t)t * K + tower) * N + envs[e];
            const float ku = cur[e].ku, kt = cur[e].kt;
            const float dhv[2] = {cur[e].dh.x + ch[e][0] * ku, cur[e].dh.y + ch[e][1] * ku}, cv[2] = {cur[e].c.x, cur[e].c.y}, cpm[2] = {cur[e].cp.x * kt, cur[e].cp.y * kt};
            const float iv[2] = {cur[e].gi.x, cur[e].gi.y}, fv[2] = {cur[e].gf.x, cur[e].gf.y}, ov[2] = {cur[e].go.x, cur[e].go.y}, gv[2] = {cur[e].gg.x, cur[e].gg.y};
            float dzv[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float tc = tanhf(cv[u]);
                const float dc = cc[e][u] * ku + dhv[u] * ov[u] * (1.f - tc * tc);
                dzv[u][0] = dc * gv[u] * iv[u] * (1.f - iv[u]); dzv[u][1] = dc * cpm[u] * fv[u] * (1.f - fv[u]);
                dzv[u][2] = dhv[u] * tc * ov[u] * (1.f - ov[u]); dzv[u][3] = dc * iv[u] * (1.f - gv[u] * gv[u]);
                cc[e][u] = dc * fv[u];
            }
            if (valid[e]) {
                float* dr = dz + row * G4 + 2 * cg;
#pragma unroll
                for (int g = 0; g < 4; ++g) { *reinterpret_cast<float2*>(dr + g * H) = make_float2(dzv[0][g], dzv[1][g]); dbacc[0][g] += dzv[0][g]; dbacc[1][g] += dzv[1][g]; }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int g = 0; g < 4; ++g) dzT[8 * cg + 4 * u + g][4 * eg + e] = dzv[u][g];
        }
        __syncthreads();
        float acc[4][2];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e][0] = acc[e][1] = 0.f;
#pragma unroll 8
        for (int col = 0; col < G4; ++col) {
            const float4 x = *reinterpret_cast<const float4*>(&dzT[col][4 * eg]);
            const float2 w = *reinterpret_cast<const float2*>(&WT[col][2 * cg]);
            acc[0][0] = fmaf(x.x, w.x, acc[0][0]); acc[0][1] = fmaf(x.x, w.y, acc[0][1]); acc[1][0] = fmaf(x.y, w.x, acc[1][0]); acc[1][1] = fmaf(x.y, w.y, acc[1][1]);
            acc[2][0] = fmaf(x.z, w.x, acc[2][0]); acc[2][1] = fmaf(x.z, w.y, acc[2][1]); acc[3][0] = fmaf(x.w, w.x, acc[3][0]); acc[3][1] = fmaf(x.w, w.y, acc[3][1]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) { ch[e][0] = acc[e][0]; ch[e][1] = acc[e][1]; cur[e] = nxt[e]; }
        __syncthreads();
    }
    if (db_part) {   // the 8 env groups of a unit pair are 8 consecutive lanes: fixed-order butterfly, one partial per CTA (summed on the host side: deterministic)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float v = dbacc[u][g];
                v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (eg == 0) db_part[((size_t)blockIdx.x * K + tower) * G4 + g * H + 2 * cg + u] = v;
            }
    }
}
void launch_lstm_seq_fwd_fma(int T, int K, int N, const float* xw, const float* wh, const float* c0, const float* h0, const float* keep, float* gates, float* Cs,
                         float* Hs, const float* bias, float* HM, cudaStream_t st) {
    dim3 grid((N + SEQ_TM - 1) / SEQ_TM, K);
    lstm_seq_fwd_kernel<<<grid, SEQ_THR, 0, st>>>(T, K, N, xw, wh, c0, h0, keep, gates, Cs, Hs, bias, HM);
}
int lstm_seq_ctas(int N) { return (N + SEQ_TM - 1) / SEQ_TM; }
void launch_lstm_seq_bwd_fma(int T, int K, int N, const float* dH, const float* wh, const float* c0, const float* keep, const float* gates, const float* Cs,
                         float* dz, float* db_part, cudaStream_t st) {
    dim3 grid((N + SEQ_TM - 1) / SEQ_TM, K);
    constexpr int smem = sizeof(float) * (G4 * H + G4 * SEQ_TM);
    {   // the attribute is per device (a process may hold envs / policies on several GPUs through the C ABI)
        static unsigned long long configured_devices = 0ull; int dev = 0; cudaGetDevice(&dev);
        if (dev >= 64 || !((configured_devices >> dev) & 1ull)) { cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); if (dev < 64) configured_devices |= 1ull << dev; }
    }
    lstm_seq_bwd_kernel<<<grid, SEQ_THR, smem, st>>>(T, K, N, dH, wh, c0, keep, gates, Cs, dz, db_part);
}

void launch_lstm_pw_fwd(int rows, int n_env, const float* z, const float* c_prev_masked, const float* keep_next, float* gates, float* c_out,
                        float* h_out, float* hm_next, float* cm_next, cudaStream_t st) {
    lstm_pw_fwd_kernel<<<(rows * H + 255) / 256, 256, 0, st>>>(rows, n_env, z, c_prev_masked, keep_next, gates, c_out, h_out, hm_next, cm_next);
}
void launch_lstm_pw_bwd(int rows, int n_env, const float* dh_out, const float* carry_h, const float* carry_c, const float* keep_up,
                        const float* gates, const float* c, const float* c_prev_masked, float* dz, float* dcm_prev, cudaStream_t st) {
    lstm_pw_bwd_kernel<<<(rows * H + 255) / 256, 256, 0, st>>>(rows, n_env, dh_out, carry_h, carry_c, keep_up, gates, c, c_prev_masked, dz, dcm_prev);
}

}  // namespace irrl
