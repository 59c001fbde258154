// Heads + PPO2 loss + their gradients in one pass over the top-layer activations (flex_gym/algo/ppo2/ppo2.py:152-175 over the heads of
// run_bp_v5.py:167-176): per sample (t, n)
//   mean = h_pi . pi_w + pi_b ; v = h_v . vf_w + vf_b ; neglogp = 0.5 sum z^2 + 6 ln(2 pi) + sum logstd , z = (a - mean) exp(-logstd)
//   ratio = exp(old_neglogp - neglogp) ; pg = max(-adv ratio, -adv clip(ratio, 1 - eps, 1 + eps))
//   vclip = v_old + clip(v - v_old, -eps, eps) ; vf = max((v - R)^2, (vclip - R)^2)
//   loss = mean(pg) + vf_coef 0.5 mean(vf)            (the entropy term depends on logstd only and stays in PyTorch)
// and, in the same thread, d loss / d mean, d loss / d v (the sub-gradient conventions of torch.maximum / torch.clamp: an exact tie splits
// evenly, the clamp passes gradient on its closed interval), pushed through the heads into dH [T,2,N,48].  Outputs per sample:
// G [T,N,16] = (d loss / d mean[12], d loss / d v, 0, 0, 0) for the head-weight gradients, dH; per CTA: partial sums of the loss terms,
// approxkl, clipfrac and d loss / d logstd.  Thread = sample: the 192-byte rows are read and written row-per-lane (32 L1 tag look-ups per
// instruction, ~1 ms for the 4.8 GB of a 8192 x 750 batch); replaces three cuBLAS fp32 products on [T,2,N,48] views and ~20 element-wise /
// reduction launches of the autograd graph.
#include <algorithm>
#include "env_device.cuh"
#include "env_kernels.h"

namespace irrl {

struct HeadLossArgs {
    const float* H1;                     // [T,2,N,48]
    const float *pi_w, *pi_b, *vf_w, *vf_b, *logstd;       // [48,12] [12] [48] [1] [12]
    const float *actions, *adv, *ret, *old_v, *old_nlp;    // [T,N,12] [T,N] ...
    float* dH;                           // [T,2,N,48]
    float* G;                            // [T,N,16]
    float* partial;                      // [gridDim.x][HL_PART]
    float cliprange, vf_coef, inv_count;
    int T, N;
};
constexpr int HL_THR = 256;
// shared-memory reads the compiler must neither hoist out of the row loop nor keep alive between the two passes (576 head weights: spills)
__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

__global__ void __launch_bounds__(HL_THR) ppo_head_loss_kernel(const __grid_constant__ HeadLossArgs A) {
    __shared__ __align__(16) float pw[48][12];
    __shared__ __align__(16) float pwT[12][48];       // second copy for the backward pass: reusing pw there makes the compiler keep all 576 loads of the forward pass alive (spills)
    __shared__ float vw[48], pb[12], ls[12], istd[12];
    __shared__ float red[HL_THR / 32][HL_PART];
    const int t_ = threadIdx.x;
    for (int i = t_; i < 48 * 12; i += HL_THR) { (&pw[0][0])[i] = A.pi_w[i]; pwT[i % 12][i / 12] = A.pi_w[i]; }
    if (t_ < 48) vw[t_] = A.vf_w[t_];
    if (t_ < 12) { pb[t_] = A.pi_b[t_]; ls[t_] = A.logstd[t_]; istd[t_] = expf(-A.logstd[t_]); }
    __syncthreads();
    const float vb = A.vf_b[0], eps = A.cliprange;
    float s_pg = 0.f, s_vf = 0.f, s_kl = 0.f, s_clip = 0.f, s_ls[12];
#pragma unroll
    for (int a = 0; a < 12; ++a) s_ls[a] = 0.f;
    float sum_ls = 0.f;
#pragma unroll
    for (int a = 0; a < 12; ++a) sum_ls += ls[a];
    const long long rows = (long long)A.T * A.N;
    for (long long row = (long long)blockIdx.x * HL_THR + t_; row < rows; row += (long long)gridDim.x * HL_THR) {
        const long long t = row / A.N, n = row - t * A.N;
        const float4* hp = reinterpret_cast<const float4*>(A.H1 + ((t * 2 + 0) * A.N + n) * 48);
        const float4* hv = reinterpret_cast<const float4*>(A.H1 + ((t * 2 + 1) * A.N + n) * 48);
        float mean[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) mean[a] = pb[a];
        float v = vb;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const float4 x = hp[c], y = hv[c];
            const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w0 = lds128(&pw[4 * c + j][0]), w1 = lds128(&pw[4 * c + j][4]), w2 = lds128(&pw[4 * c + j][8]);
                mean[0] = fmaf(xs[j], w0.x, mean[0]); mean[1] = fmaf(xs[j], w0.y, mean[1]); mean[2] = fmaf(xs[j], w0.z, mean[2]); mean[3] = fmaf(xs[j], w0.w, mean[3]);
                mean[4] = fmaf(xs[j], w1.x, mean[4]); mean[5] = fmaf(xs[j], w1.y, mean[5]); mean[6] = fmaf(xs[j], w1.z, mean[6]); mean[7] = fmaf(xs[j], w1.w, mean[7]);
                mean[8] = fmaf(xs[j], w2.x, mean[8]); mean[9] = fmaf(xs[j], w2.y, mean[9]); mean[10] = fmaf(xs[j], w2.z, mean[10]); mean[11] = fmaf(xs[j], w2.w, mean[11]);
                v = fmaf(ys[j], vw[4 * c + j], v);
            }
        }
        const float4* ap = reinterpret_cast<const float4*>(A.actions + row * 12);
        float z[12];
        {
            const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2];
            const float act[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
            for (int a = 0; a < 12; ++a) z[a] = (act[a] - mean[a]) * istd[a];
        }
        float nlp = 0.f;
#pragma unroll
        for (int a = 0; a < 12; ++a) nlp = fmaf(0.5f * z[a], z[a], nlp);
        nlp += 0.5f * 1.8378770664093453f * 12.f + sum_ls;
        const float adv = A.adv[row], R = A.ret[row], ov = A.old_v[row], onlp = A.old_nlp[row];
        const float ratio = expf(onlp - nlp);
        const float rc = fminf(fmaxf(ratio, 1.f - eps), 1.f + eps);
        const float pg1 = -adv * ratio, pg2 = -adv * rc;
        const float inside = (ratio >= 1.f - eps && ratio <= 1.f + eps) ? 1.f : 0.f;
        const float sel = pg1 > pg2 ? 1.f : (pg2 > pg1 ? inside : 0.5f + 0.5f * inside);      // d max(pg1, pg2) / d ratio = -adv * sel
        s_pg += fmaxf(pg1, pg2);
        const float dkl = nlp - onlp; s_kl = fmaf(0.5f * dkl, dkl, s_kl);
        s_clip += fabsf(ratio - 1.f) > eps ? 1.f : 0.f;
        const float g_nlp = adv * ratio * sel * A.inv_count;                                    // d ratio / d neglogp = -ratio
        const float dv = v - ov, vclip = ov + fminf(fmaxf(dv, -eps), eps);
        const float e1 = v - R, e2 = vclip - R, l1 = e1 * e1, l2 = e2 * e2;
        const float inside_v = (dv >= -eps && dv <= eps) ? 1.f : 0.f;
        s_vf += fmaxf(l1, l2);
        const float dvf = l1 > l2 ? 2.f * e1 : (l2 > l1 ? 2.f * e2 * inside_v : e1 + e2 * inside_v);
        const float g_v = A.vf_coef * 0.5f * A.inv_count * dvf;
        float gm[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) { gm[a] = -g_nlp * z[a] * istd[a]; s_ls[a] = fmaf(g_nlp, 1.f - z[a] * z[a], s_ls[a]); }
        float4* gp = reinterpret_cast<float4*>(A.G + row * 16);
        gp[0] = make_float4(gm[0], gm[1], gm[2], gm[3]); gp[1] = make_float4(gm[4], gm[5], gm[6], gm[7]); gp[2] = make_float4(gm[8], gm[9], gm[10], gm[11]);
        gp[3] = make_float4(g_v, 0.f, 0.f, 0.f);
        float4* dp = reinterpret_cast<float4*>(A.dH + ((t * 2 + 0) * A.N + n) * 48);
        float4* dq = reinterpret_cast<float4*>(A.dH + ((t * 2 + 1) * A.N + n) * 48);
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int a = 0; a < 12; ++a) {
                const float4 w = lds128(&pwT[a][4 * c]);
                o.x = fmaf(gm[a], w.x, o.x); o.y = fmaf(gm[a], w.y, o.y); o.z = fmaf(gm[a], w.z, o.z); o.w = fmaf(gm[a], w.w, o.w);
            }
            dp[c] = o;
            dq[c] = make_float4(g_v * vw[4 * c], g_v * vw[4 * c + 1], g_v * vw[4 * c + 2], g_v * vw[4 * c + 3]);
        }
    }
    // fixed-order reduction: lanes (butterfly), then warps, one partial per CTA (summed afterwards: deterministic)
    float vals[HL_PART] = {s_pg, s_vf, s_kl, s_clip, s_ls[0], s_ls[1], s_ls[2], s_ls[3], s_ls[4], s_ls[5], s_ls[6], s_ls[7], s_ls[8], s_ls[9], s_ls[10], s_ls[11]};
#pragma unroll
    for (int i = 0; i < HL_PART; ++i) {
        float x = vals[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((t_ & 31) == 0) red[t_ >> 5][i] = x;
    }
    __syncthreads();
    if (t_ < HL_PART) {
        float x = 0.f;
#pragma unroll
        for (int w = 0; w < HL_THR / 32; ++w) x += red[w][t_];
        A.partial[(size_t)blockIdx.x * HL_PART + t_] = x;
    }
}

// x *= *scale unless *scale == 1 (the upstream gradient of a terminal loss): the common case costs one launch and no memory traffic
__global__ void scale_unless_one_kernel(float4* __restrict__ x, long long n4, const float* __restrict__ scale) {
    const float s = *scale;
    if (s == 1.0f) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) { float4 v = x[i]; v.x *= s; v.y *= s; v.z *= s; v.w *= s; x[i] = v; }
}
void launch_scale_unless_one(float* x, long long n, const float* scale, cudaStream_t st) {
    scale_unless_one_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<float4*>(x), n / 4, scale);
}

int ppo_head_loss_ctas(long long rows) {
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long need = (rows + HL_THR - 1) / HL_THR;
    return (int)std::max<long long>(1, std::min<long long>(need, (long long)sms * 4));
}
void launch_ppo_head_loss(const float* H1, const float* pi_w, const float* pi_b, const float* vf_w, const float* vf_b, const float* logstd, const float* actions,
                          const float* adv, const float* ret, const float* old_v, const float* old_nlp, float* dH, float* G, float* partial,
                          float cliprange, float vf_coef, float inv_count, int T, int N, cudaStream_t st) {
    HeadLossArgs a{H1, pi_w, pi_b, vf_w, vf_b, logstd, actions, adv, ret, old_v, old_nlp, dH, G, partial, cliprange, vf_coef, inv_count, T, N};
    ppo_head_loss_kernel<<<ppo_head_loss_ctas((long long)T * N), HL_THR, 0, st>>>(a);
}

}  // namespace irrl
