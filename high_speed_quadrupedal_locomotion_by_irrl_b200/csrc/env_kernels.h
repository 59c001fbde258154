// Host-visible launchers of the env kernels (env_kernels.cu) and the LSTM / GAE kernels (policy_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "irrl_params.h"

namespace irrl {

struct StepArgs {
    EnvParams P;
    DevState S;
    const float* action;   // [N,12] device
    float* ob;             // [N,35] device (may be null)
    float* reward;         // [N]
    uint8_t* done;         // [N]
    float* extra;          // [N,6] (may be null)
    float* ep_ret_out;     // [N] episode return of the envs that finished this step (may be null)
    int* ep_len_out;       // [N]
    uint32_t tick;
    int r_begin, r_end;    // env range of this launch of the step kernel (irrl_act_step pipelines chunks of envs over several streams); whole shard = [0, N)
};

void launch_env_step(const StepArgs& a, cudaStream_t st);
void launch_env_reset(const StepArgs& a, cudaStream_t st);
void launch_env_meteor(const StepArgs& a, int respawn_only, cudaStream_t st);   // Crutial: True only; before every env step / after a reset
void launch_env_observe(const EnvParams& P, const DevState& S, float* ob, cudaStream_t st);
void launch_env_probe(const EnvParams& P, const DevState& S, float* M, float* Minv, float* h, cudaStream_t st);
void launch_env_integrate(const EnvParams& P, const DevState& S, const float* tau, float* contact_out, cudaStream_t st);
void launch_env_get_state(const EnvParams& P, const DevState& S, float* out, cudaStream_t st);
void launch_env_set_state(const EnvParams& P, const DevState& S, const float* in, cudaStream_t st);
void launch_env_init(const EnvParams& P, const DevState& S, cudaStream_t st);

// ---- policy (policy_kernels.cu)
struct PolicyWeights {      // device pointers, fp32, layouts as in the reference pkl (run_bp_v5.py:143-170)
    const float* wx[4];     // lstm_pi0, lstm_pi1, lstm_v0, lstm_v1 : [in,192]
    const float* wh[4];     // [48,192]
    const float* b[4];      // [192]
    const float* pi_w;      // [48,12]
    const float* pi_b;      // [12]
    const float* vf_w;      // [48]
    const float* vf_b;      // [1]
    const float* logstd;    // [12]
    const float* wcat[4];   // derived: [96][192] = [wx ; wh ; 0-pad] with gate-interleaved columns (unit pair, unit, gate)
    const float* bperm[4];  // derived: [192] biases in the same column order
    const unsigned char* tcblob;   // derived: 2 towers x TC_BLOB_BYTES, hi/lo-split UMMA-layout weights (policy_tc_kernels.cu)
};
constexpr int TC_BIAS_BYTES = 4096;   // [2][192] gate biases (gate-interleaved), [16] head bias, [16] logstd, [16] sigma = exp(logstd),
                                      // [48][12] head weights (12 means / 1 value, zero padded), zero pad
constexpr int TC_BLOB_BYTES = (11 + 12) * 12288 + TC_BIAS_BYTES;
struct ActArgs {
    PolicyWeights W;
    const float* obs;       // [N,35]
    const uint8_t* done;    // [N] mask = done(t-1) (may be null = zeros)
    float* state;           // [N,384] in/out
    float* action;          // [N,12] sampled, unclipped (PPO:523)
    float* clipped;         // [N,12] clip(action, -1, 1) (PPO:529-531), may be null
    float* value;           // [N]
    float* neglogp;         // [N]
    float* mean;            // [N,12] may be null
    float* obs_store;       // [N,35] copy of obs (Runner's mb_obs), may be null
    uint8_t* done_store;    // [N] copy of the mask (Runner's mb_dones), may be null
    int N; int deterministic;
    uint32_t seed, env_offset, tick;
    int n_total = 0;        // when this launch covers a chunk of a larger batch: size of the whole batch (selects the same kernel for every chunk); 0 = N
};
void launch_lstm_act(const ActArgs& a, cudaStream_t st);          // dispatches on N and g_act_path
void launch_lstm_act_fma(const ActArgs& a, cudaStream_t st);      // fp32 FMA kernel (policy_kernels.cu)
void launch_lstm_act_tc(const ActArgs& a, cudaStream_t st);       // tcgen05 kernel (policy_tc_kernels.cu)
extern int g_act_path;                                             // 0 auto (tensor cores from 256 envs), 1 FMA, 2 tensor cores
int launch_tc_gemm_probe(const float* dA, const float* dB, float* dD, int K, int N, int variant, cudaStream_t st);
void tc_timeline(int enable, long long* out16);
int launch_tc_mma_rate(int N, int reps, unsigned layout_type, unsigned lbo, unsigned sbo, unsigned kadv, long long* d_out, cudaStream_t st);
void pack_tc_blob(const float* wx0, const float* wh0, const float* b0, const float* wx1, const float* wh1, const float* b1, const float* head_w, const float* head_b,
                  int head_cols, const float* logstd, unsigned char* out);
void launch_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values, const uint8_t* last_dones,
                float* adv, float* ret, int T, int N, float gamma, float lam, cudaStream_t st);

void launch_lstm_seq_fwd(int T, int K, int N, const float* xw, const float* wh, const float* c0, const float* h0, const float* keep, float* gates, float* Cs,
                         float* Hs, const float* bias, float* HM, cudaStream_t st);
void launch_lstm_seq_bwd(int T, int K, int N, const float* dH, const float* wh, const float* c0, const float* keep, const float* gates, const float* Cs,
                         float* dz, float* db_part, cudaStream_t st);
int lstm_seq_ctas(int N);   // CTAs per tower of the sequence kernels = rows of db_part
int lstm_seq_set_path(int path);   // 0 tensor-core kernels (default), 1 FP32-FMA kernels; returns the previous value
// the FP32-FMA variants (policy_kernels.cu); launch_lstm_seq_fwd / _bwd (lstm_seq_mma.cu) dispatch to the tensor-core kernels unless IRRL_SEQ_PATH=fma
void launch_lstm_seq_fwd_fma(int T, int K, int N, const float* xw, const float* wh, const float* c0, const float* h0, const float* keep, float* gates, float* Cs,
                             float* Hs, const float* bias, float* HM, cudaStream_t st);
void launch_lstm_seq_bwd_fma(int T, int K, int N, const float* dH, const float* wh, const float* c0, const float* keep, const float* gates, const float* Cs,
                             float* dz, float* db_part, cudaStream_t st);
// non-recurrent products of the learner on the tensor cores (learner_gemm.cu)
int launch_proj_rows(const float* X, int x_cols, int x_has_tower, const float* W, int w_trans, float* Y, int n_out, int T, int K, int N, cudaStream_t st);
int gram_rows_ctas(int T, int N, int K);
// tcgen05 forward projections (learner_tc.cu): returns -1 when the shape / path selection asks for the warp-level kernel instead
int launch_proj_rows_tc(const float* X, int x_cols, int x_has_tower, const float* W, float* Y, int T, int K, int N, cudaStream_t st);
int proj_rows_set_path(int path);
int launch_projT_rows_tc(const float* D, const float* W, float* Y, int T, int K, int N, cudaStream_t st);
int gram2_rows_tc_ctas(int T, int N, int K);
int launch_gram2_rows_tc(const float* X, int x_cols, int x_has_tower, const float* Hs, const float* h0, const float* keep, const float* D, float* partial, int T, int K, int N,
                         cudaStream_t st);
// proj_rows_set_path: 0 tcgen05 (default), 1 warp-level MMAs; returns the previous value
int launch_gram_rows(const float* X, int x_cols, int x_has_tower, const float* D, float* partial, int T, int K, int N, cudaStream_t st);
// heads + PPO loss + gradients in one pass (ppo_head_loss.cu); partial [ppo_head_loss_ctas][HL_PART] = per-CTA sums of
// (pg, vf, 0.5 (neglogp - old)^2, clipped count, d loss / d logstd[12])
constexpr int HL_PART = 16;
int ppo_head_loss_ctas(long long rows);
void launch_scale_unless_one(float* x, long long n, const float* scale, cudaStream_t st);   // n a multiple of 4, x 16-byte aligned
void launch_ppo_head_loss(const float* H1, const float* pi_w, const float* pi_b, const float* vf_w, const float* vf_b, const float* logstd, const float* actions,
                          const float* adv, const float* ret, const float* old_v, const float* old_nlp, float* dH, float* G, float* partial,
                          float cliprange, float vf_coef, float inv_count, int T, int N, cudaStream_t st);
void launch_lstm_pw_fwd(int rows, int n_env, const float* z, const float* c_prev_masked, const float* keep_next, float* gates, float* c_out,
                        float* h_out, float* hm_next, float* cm_next, cudaStream_t st);
void launch_lstm_pw_bwd(int rows, int n_env, const float* dh_out, const float* carry_h, const float* carry_c, const float* keep_up,
                        const float* gates, const float* c, const float* c_prev_masked, float* dz, float* dcm_prev, cudaStream_t st);

}  // namespace irrl
