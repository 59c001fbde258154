/* irrl_b200.h -- C ABI of the B200-native bp5 vectorised environment + LSTM act path.
 *
 * This is the drop-in boundary for the reference's pybind11 module `_flexible_robot` (class `FlexibleGymEnv`,
 * flex_gym/env/raisim_gym.cpp:14-47) which binds VectorizedEnvironment<ENVIRONMENT>
 * (flex_gym/env/VectorizedEnvironment.hpp:127-382).  Every entry point below names the reference method it
 * replaces.  Plain pointers and sizes only; no torch / pybind types.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure; irrl_last_error() returns the message
 *     (the reference aborts the process via RSFATAL instead: RaisimGymEnv.hpp:41-42, VectorizedEnvironment.hpp:186).
 *   - data pointers may be HOST or DEVICE memory (detected with cudaPointerGetAttributes).  Host buffers are staged
 *     through pinned memory and the call returns after the results have landed in the caller's buffer, i.e. the
 *     in-place semantics of the reference's Eigen::Ref arguments (RaisimGymVecEnv.py:26-52).  Device buffers are used
 *     in place on the env's stream and the call returns without synchronising.
 *   - 2-D arrays are C-contiguous row-major float32 (EigenRowMajorMat, RaisimGymEnv.hpp:46-49); `done` is one byte
 *     per env (EigenBoolVec).
 *   - N = num_envs of THIS handle (a shard when env_offset/num_envs select a slice of a larger job).
 */
#ifndef IRRL_B200_H
#define IRRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct irrl_env irrl_env;
typedef struct irrl_policy irrl_policy;

#define IRRL_OB_DIM 35        /* Environment.hpp:360 */
#define IRRL_ACTION_DIM 12    /* Environment.hpp:361 */
#define IRRL_EXTRA_DIM 6      /* Environment.hpp:944-949 */
#define IRRL_ORIGIN_STATE_DIM 41 /* Environment.hpp:1330-1334 : gc(19) gv(18) contact(4) */
#define IRRL_STATE_DIM 192    /* flat state vector, see irrl_get_state */
#define IRRL_LSTM_STATE_DIM 384 /* run_bp_v5.py:136-137 */

const char* irrl_last_error(void);
const char* irrl_version(void);

/* ---- construction: FlexibleGymEnv(resourceDir, cfgYaml)  raisim_gym.cpp:16, VectorizedEnvironment.hpp:132-138.
 * cfg_yaml is the dumped `environment:` map (run_bp_v5.py:205-207); every key of Environment.hpp:1594-1659 and
 * VectorizedEnvironment.hpp:146-171 is mandatory.  device = CUDA ordinal.  env_offset = global id of env 0 of this
 * handle (keys the counter-based RNG so results do not depend on how a job is sharded over GPUs). */
int irrl_create(const char* resource_dir, const char* cfg_yaml, int device, int env_offset, irrl_env** out);
void irrl_destroy(irrl_env* env);
/* VectorizedEnvironment::init  VectorizedEnvironment.hpp:145-194 (allocates state, samples domain randomisation, resets) */
int irrl_init(irrl_env* env);
int irrl_set_stream(irrl_env* env, void* cuda_stream);

/* ---- dimensions  VectorizedEnvironment.hpp:336-342, 196 */
int irrl_get_ob_dim(irrl_env* env);
int irrl_get_action_dim(irrl_env* env);
int irrl_get_extra_info_dim(irrl_env* env);
int irrl_get_num_envs(irrl_env* env);
const char* irrl_get_extra_info_name(irrl_env* env, int index);
int irrl_get_origin_state_dim(irrl_env* env);   /* Environment.hpp:1330 */

/* ---- the hot path */
/* VectorizedEnvironment::reset(ob[N,35])  VectorizedEnvironment.hpp:201-207 */
int irrl_reset(irrl_env* env, float* ob);
/* VectorizedEnvironment::observe(ob[N,35])  VectorizedEnvironment.hpp:209-212 */
int irrl_observe(irrl_env* env, float* ob);
/* VectorizedEnvironment::step(action[N,12], ob[N,35], reward[N], done[N], extraInfo[N,6])  VectorizedEnvironment.hpp:268-278 */
int irrl_step(irrl_env* env, const float* action, float* ob, float* reward, uint8_t* done, float* extra_info);
/* VectorizedEnvironment::testStep: env 0 only  VectorizedEnvironment.hpp:280-290 */
int irrl_test_step(irrl_env* env, const float* action, float* ob, float* reward, uint8_t* done, float* extra_info);
/* episode bookkeeping of RaisimGymVecEnv.step (RaisimGymVecEnv.py:42-50) done on the device: for every env whose
 * `done` was set by the last irrl_step, the episode return and length; zeros elsewhere. */
int irrl_last_episode_stats(irrl_env* env, float* ep_return, int32_t* ep_length);
/* running (unfinished) episode return/length, RaisimGymVecEnv._update_epi_info (RaisimGymVecEnv.py:103-113); clear != 0 resets them */
int irrl_running_episode_stats(irrl_env* env, float* ep_return, int32_t* ep_length, int clear);

/* ---- control plane mirrors */
int irrl_set_seed(irrl_env* env, int seed);                          /* VectorizedEnvironment.hpp:308-312 */
int irrl_close(irrl_env* env);                                       /* VectorizedEnvironment.hpp:314 */
int irrl_is_terminal_state(irrl_env* env, uint8_t* terminal);        /* VectorizedEnvironment.hpp:319-324 */
int irrl_set_simulation_time_step(irrl_env* env, double dt);         /* VectorizedEnvironment.hpp:326 */
int irrl_set_control_time_step(irrl_env* env, double dt);            /* VectorizedEnvironment.hpp:331 */
int irrl_curriculum_update(irrl_env* env);                           /* VectorizedEnvironment.hpp:345 (no-op in ENV) */
int irrl_start_recording_video(irrl_env* env, const char* file);     /* VectorizedEnvironment.hpp:292 (no display: no-op) */
int irrl_stop_recording_video(irrl_env* env);                        /* VectorizedEnvironment.hpp:296 */
int irrl_show_window(irrl_env* env);                                 /* VectorizedEnvironment.hpp:300 */
int irrl_hide_window(irrl_env* env);                                 /* VectorizedEnvironment.hpp:304 */

/* ---- probes */
int irrl_origin_state(irrl_env* env, float* out /*[N,41]*/);         /* Environment.hpp:1317-1325 */
int irrl_reference_state(irrl_env* env, float* out /*[N,24]*/);      /* Environment.hpp:1339-1345 (jointRef, jointDotRef) */
int irrl_get_joint_effort(irrl_env* env, float* out /*[N,12]*/);     /* Environment.hpp:1350-1358 */
int irrl_get_generalized_force(irrl_env* env, float* out /*[N,18]*/);/* Environment.hpp:1363-1370 */
int irrl_get_inverse_mass_matrix(irrl_env* env, float* out /*[N,324]*/); /* Environment.hpp:1375-1391 (column-major) */
int irrl_get_nonlinear(irrl_env* env, float* out /*[N,18]*/);        /* Environment.hpp:1396-1402 */
int irrl_get_mass_matrix(irrl_env* env, float* out /*[N,324]*/);     /* RaiSim getMassMatrix (parity probe, north_star) */
int irrl_set_contact_coefficient(irrl_env* env, const float* coeff /*[N,3] mu, restitution, threshold*/); /* Environment.hpp:1407-1418 */
int irrl_get_sphere_info(irrl_env* env, float* out /*[N,4]*/);       /* Environment.hpp:1423-1436 (needs Crutial; returns -3 otherwise) */
/* Crutial: True (meteor sphere, Environment.hpp:717-741, 815-861): test access to its state, host memory [N,9] = position(3),
 * velocity(3), mode (0 just placed / 1 falling), radius, mass.  -3 when Crutial is False. */
int irrl_get_meteor(irrl_env* env, float* out /*[N,9]*/);
int irrl_set_meteor(irrl_env* env, const float* in /*[N,9]*/);
/* The robot description of the resource directory (Environment.hpp:231 addArticulatedSystem(resourceDir + "/black_panther.urdf")).
 * irrl_create reads <resource_dir>/black_panther.urdf (or <resource_dir>/urdf/black_panther.urdf) when present and fails if its four
 * legs are not mirror images (the kernels carry one set of leg constants); without the file the built-in constants of the shipped
 * description are used.  irrl_parse_urdf needs no GPU: 40 floats = I0(3) I1(3) I2(xx,yy,zz,|yz|) I3(3) rotor(3) off1x off1y off2y
 * toe_z toe_r box_half(3) joint_damping m0 com0(3) m1 com1(3) m2 com2(3) m3 com3z knee_z; path NULL returns the built-in model. */
int irrl_parse_urdf(const char* path, float* out40);
int irrl_get_model_params(irrl_env* env, float* out /*[N,94]: mu,rest,thr, 13 x (mass, com3, joint offset3)*/);

/* ---- state injection / extraction (the reference only has setState internally, Environment.hpp:618-622).
 * Flat layout per env (float32[192]):
 *   [0,19) gc  [19,37) gv  [37,49) pTarget12Last  [49,61) torque_last  [61,64) command  [64,67) command_filtered
 *   [67,79) jointRef  [79,91) jointDotRef  [91,103) EndEffectorRef  103 t0  104 frame_idx  [105,109) contact
 *   [109,144) obDouble  [144,179) obDouble_last  [179,191) applied joint torque  191 itera
 * current_time_ of the reference = t0 + frame_idx * control_dt. */
int irrl_get_state(irrl_env* env, float* out /*[N,192]*/);
int irrl_set_state(irrl_env* env, const float* in /*[N,192]*/);
int irrl_set_tick(irrl_env* env, uint32_t tick);      /* RNG step counter */
uint32_t irrl_get_tick(irrl_env* env);
/* one world.integrate() (Environment.hpp:768) with given joint torques; contact_out[N,16] = per foot (active, impulse xyz) */
int irrl_integrate(irrl_env* env, const float* tau /*[N,12]*/, float* contact_out);
int irrl_get_solver_sweeps(irrl_env* env, int32_t* out /*[N]*/);
/* reference-trajectory table mode (ManualTraj False): ref[rows,30] float32  VectorizedEnvironment.hpp:158-182, Environment.hpp:17-21 */
int irrl_set_ref_traj(irrl_env* env, const float* table, int rows);

/* ---- heightfield terrain (Terrain: True; the reference adds a RaiSim HeightMap, Environment.hpp:252-265).  heights[nx][ny] float32,
 * grid of x_size x y_size metres centred on (cx, cy), cells split along the (i,j)-(i+1,j+1) diagonal.  With Terrain: True and no table
 * supplied before irrl_init, the table of YAML key `terrain_kind` ("perlin" (default) | "stairs" | "flat"; `stair_rise`, `stair_run`,
 * `stair_start`, `terrain_seed`) is generated with the reference's grid (5000 x 500 samples over 500 x 20 m). */
int irrl_set_heightfield(irrl_env* env, const float* heights, int nx, int ny, double x_size, double y_size, double cx, double cy);
int irrl_generate_terrain(irrl_env* env, const char* kind, int nx, int ny, double x_size, double y_size);
int irrl_get_heightfield(irrl_env* env, float* heights, int* nx, int* ny, double* x_size, double* y_size, double* cx, double* cy);

/* ---- LSTM policy act  (CustomLSTMPolicy.step  run_bp_v5.py:178-185)
 * params: the 19 arrays of the reference pkl in tf.trainable_variables() order, concatenated:
 * lstm_pi0{wx[35,192],wh[48,192],b[192]} lstm_pi1{wx[48,192],wh,b} lstm_v0{...} lstm_v1{...} vf{w[48,1],b[1]}
 * pi{w[48,12],b[12]} pi/logstd[12] q{w[48,12],b[12]}  = 70741 floats. */
#define IRRL_POLICY_NUM_PARAMS 70741
int irrl_policy_create(int device, const float* params, irrl_policy** out);
int irrl_policy_set_params(irrl_policy* pol, const float* params);
void irrl_policy_destroy(irrl_policy* pol);
/* Which act kernel irrl_policy_act / irrl_rollout launch: 0 = automatic (tcgen05 tensor-core kernel from 256 environments,
 * fp32 FMA kernel below), 1 = always the fp32 FMA kernel, 2 = always the tcgen05 kernel.  Both follow run_bp_v5.py:143-176. */
int irrl_policy_set_act_path(int mode);
/* obs[N,35], done[N] (mask = done of the previous step, may be NULL), state[N,384] in/out, action[N,12] (unclipped),
 * clipped[N,12] (may be NULL), value[N], neglogp[N]; deterministic != 0 -> action = mean.  seed/env_offset/tick key the
 * Gaussian draws. */
int irrl_policy_act(irrl_policy* pol, void* cuda_stream, int n, const float* obs, const uint8_t* done, float* state,
                    float* action, float* clipped, float* value, float* neglogp, int deterministic,
                    uint32_t seed, uint32_t env_offset, uint32_t tick);
/* Runner.run (ppo2.py:519-552) on the device: T x [act -> clip -> env step], storing the rollout in DEVICE buffers
 * laid out [T,N,...]; obs0/state/done0 carry the runner state in and out.  All pointers must be device memory. */
typedef struct irrl_rollout_buffers {
    float* obs;        /* [T,N,35] */
    float* actions;    /* [T,N,12] unclipped (ppo2.py:523) */
    float* values;     /* [T,N] */
    float* neglogps;   /* [T,N] */
    float* rewards;    /* [T,N] */
    uint8_t* dones;    /* [T,N] done flag BEFORE step t (mb_dones, ppo2.py:526) */
    float* cur_obs;    /* [N,35] in: obs before the first step; out: obs after the last */
    uint8_t* cur_done; /* [N]    in/out */
    float* state;      /* [N,384] in/out */
    float* ep_return;  /* [T,N] episode return where an episode ended at step t else 0 (may be NULL) */
    int32_t* ep_length;/* [T,N] (may be NULL) */
} irrl_rollout_buffers;
int irrl_rollout(irrl_env* env, irrl_policy* pol, int T, const irrl_rollout_buffers* buf, int deterministic);
/* One host entry per control step of a rollout: model.step (run_bp_v5.py:178-185) -> np.clip (ppo2.py:529-531) -> env.step
 * (RaisimGymVecEnv.py:26-52 -> VectorizedEnvironment.hpp:268-278), i.e. what Runner.run does between two iterations of its loop
 * (ppo2.py:519-538) with ONE call instead of two.  Buffers are either all device memory (no copies, asynchronous on the env's stream)
 * or all PAGE-LOCKED host memory (irrl_host_register): then the environments are cut into `chunks` pieces (0 = automatic) that
 * travel through separate streams, so that the copies of one chunk overlap the kernels of another, and the call returns when
 * everything has landed.  obs / done may alias next_obs / next_done (the outputs of one call are the inputs of the next).  The LSTM
 * state is device memory (the Runner only threads `states` through, ppo2.py:520).  act_tick keys the Gaussian draws. */
typedef struct irrl_act_step_io {
    const float* obs;        /* [N,35] in  */
    const uint8_t* done;     /* [N]    in: done flag of the previous step = LSTM mask (may be NULL) */
    float* state;            /* [N,384] in/out, DEVICE */
    float* action;           /* [N,12] out: sampled, unclipped (mb_actions) */
    float* clipped;          /* [N,12] out: what the env received (may be NULL) */
    float* value;            /* [N]    out */
    float* neglogp;          /* [N]    out */
    float* next_obs;         /* [N,35] out */
    float* reward;           /* [N]    out */
    uint8_t* next_done;      /* [N]    out */
    float* extra;            /* [N,6]  out (may be NULL) */
} irrl_act_step_io;
int irrl_act_step(irrl_env* env, irrl_policy* pol, const irrl_act_step_io* io, int deterministic, uint32_t act_tick, int chunks);
/* sequence-persistent BPTT through one LSTM layer of K towers (the PPO learner's recurrence, ppo2.py:136-197 over run_bp_v5.py:143-166):
 * one launch per direction, a CTA owns 32 envs for all T steps.  Time-major device tensors: xw / gates / dz [T,K,N,192] (gate order
 * i,f,o,g), Cs / Hs / dH [T,K,N,48], keep [T,N] = 1 - done mask, c0 / h0 [K,N,48] (state before step 0, unmasked), wh [K,48,192]. */
int irrl_lstm_seq_fwd(void* cuda_stream, int T, int K, int n_env, const float* xw, const float* wh, const float* c0, const float* h0, const float* keep,
                      float* gates, float* Cs, float* Hs, const float* bias /*[K,192] added to xw, may be NULL*/,
                      float* HM /*[T,K,N,48] out: masked hidden state fed INTO every step (operand of dW_h), may be NULL*/);
int irrl_lstm_seq_bwd(void* cuda_stream, int T, int K, int n_env, const float* dH, const float* wh, const float* c0, const float* keep, const float* gates,
                      const float* Cs, float* dz, float* db_part /*[irrl_lstm_seq_ctas(n_env),K,192] out: per-CTA sums of dz (bias gradient), may be NULL*/);
int irrl_lstm_seq_ctas(int n_env);
/* which kernels serve irrl_lstm_seq_fwd / _bwd: 0 = tensor-core recurrence (default; forward product as a 2-term fp16 split on the f16 MMA -- it assumes
 * |W_h| < 6.5e4, the fp16 range, and |h| < 1 which the cell guarantees -- backward product as tf32 in 3xTF32 form), 1 = FP32-FMA recurrence
 * (the regression reference; also IRRL_SEQ_PATH=fma in the environment).  Returns the previous value; any other argument only queries. */
int irrl_lstm_seq_set_path(int path);
/* the non-recurrent products of the same BPTT on the tensor cores (3xTF32, fp32-grade), device pointers, time-major rows (t, tower k, env n):
 *   irrl_proj_rows  Y[T,K,N,n_out] = X . B_k, B_k = W_k [x_cols,n_out] or, with w_trans, W_k^T of W_k [n_out,x_cols]; X is [T,K,N,x_cols] or, with
 *                   x_has_tower = 0, [T,N,x_cols] shared by the towers.  Shapes served: (x_cols <= 40 or 48, n_out 192) = the input projections
 *                   x W_x (run_bp_v5.py:151-166), (x_cols 192, n_out 48) = their input gradient dz W_x^T.
 *   irrl_gram_rows  partial[irrl_gram_rows_ctas(T,K,N), K, 48, 192] = per-CTA sums of X[row,:]^T D[row,:] over the rows (t, n) of tower k, x_cols <= 48
 *                   (rows of the result past x_cols are zero); the caller sums the partials (fixed order: deterministic) = dW_x / dW_h. */
int irrl_proj_rows(void* cuda_stream, int T, int K, int n_env, const float* X, int x_cols, int x_has_tower, const float* W, int w_trans, float* Y, int n_out);
/* forward projections (n_out 192, no transpose): 0 = tcgen05 kernel (TMEM accumulators, default), 1 = warp-level MMA kernel; returns the previous value */
int irrl_proj_rows_set_path(int path);
int irrl_gram_rows_ctas(int T, int K, int n_env);
/* both weight gradients of one LSTM layer in one pass over dz, on tcgen05 (accumulator in tensor memory over all rows of a CTA):
 * partial[irrl_gram2_rows_ctas(T,K,N), K, 128, 192], rows 0..x_cols-1 = per-CTA sums of X^T D (dW_x), rows 48..95 = sums of HM^T D (dW_h) where
 * HM(t) = Hs(t-1) * keep(t) (h0 * keep(0) at t = 0) is the masked state fed into step t, formed on the fly from the layer's output sequence
 * Hs [T,K,N,48], its initial state h0 [K,N,48] and keep [T,N]; rows 96..127 are padding and NOT written.  X as in irrl_gram_rows (x_cols <= 48),
 * D [T,K,N,192].  The rows stream through the bulk-copy engine: n_env % 4 == 0 and 16-byte aligned tensors are required (-3 otherwise: use
 * irrl_gram_rows). */
int irrl_gram2_rows_ctas(int T, int K, int n_env);
int irrl_gram2_rows(void* cuda_stream, int T, int K, int n_env, const float* X, int x_cols, int x_has_tower, const float* Hs, const float* h0, const float* keep, const float* D,
                    float* partial);
int irrl_gram_rows(void* cuda_stream, int T, int K, int n_env, const float* X, int x_cols, int x_has_tower, const float* D, float* partial);
/* heads + PPO2 loss + their gradients in one pass over the top-layer activations H1 [T,2,N,48] (ppo2.py:152-175 over run_bp_v5.py:167-176):
 * per sample the Gaussian mean (pi head), the value (V head), neglogp, the clipped surrogate and the clipped value loss, and the gradients of
 *   loss = mean(pg) + vf_coef * 0.5 * mean(vf)    (scaled by inv_count = 1 / samples)
 * w.r.t. the head outputs -> G [T,N,16] = (d/d mean[12], d/d v, 0, 0, 0) and, through the heads, dH [T,2,N,48].  partial
 * [irrl_ppo_head_loss_ctas(T,N), 16] = per-CTA sums of (pg, vf, 0.5 (neglogp - old)^2, clipped count, d loss / d logstd[12]); the caller sums them
 * (fixed order: deterministic).  adv is the normalised advantage; logstd [12]; device pointers. */
int irrl_ppo_head_loss_ctas(int T, int n_env);
/* x[0:n] *= *scale on the device unless *scale == 1 (the upstream gradient of a terminal loss: no memory traffic then); n % 4 == 0, x 16-byte aligned */
int irrl_scale_unless_one(void* cuda_stream, float* x, long long n, const float* scale);
int irrl_ppo_head_loss(void* cuda_stream, int T, int n_env, const float* H1, const float* pi_w, const float* pi_b, const float* vf_w, const float* vf_b, const float* logstd,
                       const float* actions, const float* adv, const float* returns, const float* old_values, const float* old_neglogp,
                       float cliprange, float vf_coef, float inv_count, float* dH, float* G, float* partial);
/* fused element-wise halves of one LSTM training step (forward / backward through the cell), device pointers only; rows = towers * envs,
 * z / gates [rows,192] in gate order i,f,o,g, the rest [rows,48]; keep = 1 - done mask per env (run_bp_v5.py:151-153 lstm(..., masks, ...)) */
int irrl_lstm_pw_fwd(void* cuda_stream, int rows, int n_env, const float* z, const float* c_prev_masked, const float* keep_next, float* gates,
                     float* c_out, float* h_out, float* hm_next, float* cm_next);
int irrl_lstm_pw_bwd(void* cuda_stream, int rows, int n_env, const float* dh_out, const float* carry_h, const float* carry_c, const float* keep_up,
                     const float* gates, const float* c, const float* c_prev_masked, float* dz, float* dcm_prev);
/* page-lock a caller-owned host buffer (cudaHostRegister) so that irrl_step / irrl_policy_act DMA straight into it
 * instead of staging through the library's pinned buffer; RaisimGymVecEnv does this once for the buffers it owns
 * (RaisimGymVecEnv.py:14-19). */
int irrl_host_register(void* ptr, size_t bytes);
int irrl_host_unregister(void* ptr);
/* optional per-kernel timing of irrl_rollout with CUDA events on the env's stream (bench.py's roofline leg) */
int irrl_set_profiling(irrl_env* env, int on);
int irrl_get_profile(irrl_env* env, double* act_ms_total, double* step_ms_total, int64_t* launches_each);
/* register-resident FMA micro-benchmark: the non-tensor FP32 peak of this device in TFLOP/s (roofline denominator) */
int irrl_measure_fp32_peak(int device, double* tflops);
/* GAE(gamma, lambda) reverse scan  ppo2.py:554-568; all device pointers, [T,N] */
int irrl_gae(void* cuda_stream, int T, int n, const float* rewards, const float* values, const uint8_t* dones,
             const float* last_values, const uint8_t* last_dones, float gamma, float lam, float* adv, float* returns);

#ifdef __cplusplus
}
#endif
#endif /* IRRL_B200_H */
