// Fused two-tower 2x48 LSTM policy act() and GAE for sm_100a.
//
// Replaces CustomLSTMPolicy.step (run_bp_v5.py:178-185 over the graph run_bp_v5.py:143-176: stable-baselines
// `lstm`, `linear`, DiagGaussian sample + neglogp) and Runner's GAE loop (ppo2.py:554-568).
// One launch = both towers x both layers + pi/V heads + Gaussian sample + neglogp + done-mask state reset.
//
// v1 mapping (fp32 CUDA cores, exact against the numpy oracle): a CTA owns a tile of TM environments; the first
// 192 threads run the pi tower (one gate column each), the next 192 the V tower.  Inputs/hidden states of the
// tile sit transposed in shared memory ([k][env]) so a gate column reads 4 envs per LDS.128 broadcast and
// streams its weight column from L2 (the 283 KB of weights stay L2-resident across CTAs).
#include "env_device.cuh"
#include "env_kernels.h"

namespace irrl {

constexpr int TM = 16;          // environments per CTA
constexpr int H = LSTM_H;       // 48
constexpr int G4 = 4 * H;       // 192 gate columns per layer

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// gates[col] for TM envs: acc[e] = b[col] + sum_k xT[k][e] * Wx[k][col] + sum_k hT[k][e] * Wh[k][col]
template <int KIN>
__device__ __forceinline__ void gate_column(const float* __restrict__ Wx, const float* __restrict__ Wh, const float* __restrict__ bias,
                                            const float (*xT)[TM], const float (*hT)[TM], int col, float acc[TM]) {
    float b = bias[col];
#pragma unroll
    for (int e = 0; e < TM; ++e) acc[e] = b;
#pragma unroll 4
    for (int k = 0; k < KIN; ++k) {
        float w = __ldg(Wx + (size_t)k * G4 + col);
        const float4* xr = reinterpret_cast<const float4*>(xT[k]);
#pragma unroll
        for (int e4 = 0; e4 < TM / 4; ++e4) {
            float4 x = xr[e4];
            acc[4 * e4 + 0] = fmaf(x.x, w, acc[4 * e4 + 0]); acc[4 * e4 + 1] = fmaf(x.y, w, acc[4 * e4 + 1]);
            acc[4 * e4 + 2] = fmaf(x.z, w, acc[4 * e4 + 2]); acc[4 * e4 + 3] = fmaf(x.w, w, acc[4 * e4 + 3]);
        }
    }
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
        float w = __ldg(Wh + (size_t)k * G4 + col);
        const float4* hr = reinterpret_cast<const float4*>(hT[k]);
#pragma unroll
        for (int e4 = 0; e4 < TM / 4; ++e4) {
            float4 x = hr[e4];
            acc[4 * e4 + 0] = fmaf(x.x, w, acc[4 * e4 + 0]); acc[4 * e4 + 1] = fmaf(x.y, w, acc[4 * e4 + 1]);
            acc[4 * e4 + 2] = fmaf(x.z, w, acc[4 * e4 + 2]); acc[4 * e4 + 3] = fmaf(x.w, w, acc[4 * e4 + 3]);
        }
    }
}

struct __align__(16) ActSmem {
    float obsT[OB_DIM + 1][TM];        // input, transposed
    float hT[2][2][H][TM];             // [tower][layer] hidden state h(t-1) (masked), transposed
    float cS[2][2][H][TM];             // [tower][layer] cell state
    float gates[2][G4][TM];            // [tower] gate pre-activations of the current layer
    float hnew[2][H][TM];              // [tower] output of the current layer (input of the next)
    float mean[TM][ACT_DIM];
    float nlp[TM][ACT_DIM];
};

__global__ void __launch_bounds__(2 * G4) lstm_act_kernel(const __grid_constant__ ActArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ActSmem& s = *reinterpret_cast<ActSmem*>(smem_raw);
    const int t = threadIdx.x, tower = t / G4, col = t % G4;
    const int e0 = blockIdx.x * TM;
    const PolicyWeights& W = A.W;
    // ---- stage obs and (masked) state, transposed
    for (int i = t; i < TM * OB_DIM; i += 2 * G4) {
        int e = i / OB_DIM, k = i % OB_DIM; int env = min(e0 + e, A.N - 1);
        s.obsT[k][e] = A.obs[(size_t)env * OB_DIM + k];
    }
    for (int i = t; i < TM * LSTM_STATE; i += 2 * G4) {
        int e = i / LSTM_STATE, k = i % LSTM_STATE; int env = min(e0 + e, A.N - 1);
        float keep = (A.done && A.done[env]) ? 0.f : 1.f;                 // SB lstm(): c *= 1-m, h *= 1-m
        float v = A.state[(size_t)env * LSTM_STATE + k] * keep;
        int tw = k / (4 * H), rem = k % (4 * H), layer = rem / (2 * H), ch = (rem % (2 * H)) / H, u = rem % H;   // [c0,h0,c1,h1] per tower
        if (ch == 0) s.cS[tw][layer][u][e] = v; else s.hT[tw][layer][u][e] = v;
    }
    __syncthreads();
    // ---- two layers per tower
    for (int layer = 0; layer < 2; ++layer) {
        float acc[TM];
        int wi = tower * 2 + layer;
        if (layer == 0) gate_column<OB_DIM>(W.wx[wi], W.wh[wi], W.b[wi], s.obsT, s.hT[tower][0], col, acc);
        else gate_column<H>(W.wx[wi], W.wh[wi], W.b[wi], s.hnew[tower], s.hT[tower][1], col, acc);
#pragma unroll
        for (int e = 0; e < TM; ++e) s.gates[tower][col][e] = acc[e];
        __syncthreads();
        // cell update: (env, unit) pairs, gate order i,f,o,g (CustomerLstmNN.py:119-126)
        for (int i = col; i < TM * H; i += G4) {
            int u = i / TM, e = i % TM;
            float ig = sigmoidf_(s.gates[tower][u][e]), fg = sigmoidf_(s.gates[tower][H + u][e]);
            float og = sigmoidf_(s.gates[tower][2 * H + u][e]), gg = tanhf(s.gates[tower][3 * H + u][e]);
            float c = fg * s.cS[tower][layer][u][e] + ig * gg;
            float h = og * tanhf(c);
            s.hnew[tower][u][e] = h;
            int env = e0 + e;
            if (env < A.N) {
                float* st = A.state + (size_t)env * LSTM_STATE + tower * 4 * H + layer * 2 * H;
                st[u] = c; st[H + u] = h;
            }
        }
        __syncthreads();
    }
    // ---- heads
    if (tower == 0) {
        // pi: 48 -> 12, one (env, action) per thread; Gaussian sample + neglogp terms
        int e = col / ACT_DIM, a = col % ACT_DIM; int env = e0 + e;
        float m = W.pi_b[a];
#pragma unroll 8
        for (int k = 0; k < H; ++k) m = fmaf(s.hnew[0][k][e], __ldg(W.pi_w + k * ACT_DIM + a), m);
        float ls = W.logstd[a], sd = expf(ls);
        float eps = 0.f;
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
    } else if (col < TM) {
        int e = col, env = e0 + e;
        float v = W.vf_b[0];
#pragma unroll 8
        for (int k = 0; k < H; ++k) v = fmaf(s.hnew[1][k][e], __ldg(W.vf_w + k), v);
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

void launch_lstm_act(const ActArgs& a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(lstm_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ActSmem)); configured = true; }
    int grid = (a.N + TM - 1) / TM;
    lstm_act_kernel<<<grid, 2 * G4, sizeof(ActSmem), st>>>(a);
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

}  // namespace irrl
