// C entry points over the CPU oracle (bp5_oracle.hpp) for ctypes.  TEST INFRASTRUCTURE ONLY:
// loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include "bp5_oracle.hpp"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>

using namespace bp5o;

namespace {
struct Handle {
    int precision;  // 0 = double (like ENV's Eigen::VectorXd), 1 = float
    VecEnv<double> d;
    VecEnv<float> f;
    std::string err;
};
constexpr int STATE_DIM = 192;

template <typename T> void get_state(const Env<T>& e, double* s) {
    int o = 0;
    for (int i = 0; i < NQ; ++i) s[o++] = e.gc_[i];
    for (int i = 0; i < NV; ++i) s[o++] = e.gv_[i];
    for (int i = 0; i < NJ; ++i) s[o++] = e.pTarget12Last_[i];
    for (int i = 0; i < NJ; ++i) s[o++] = e.torque_last[i];
    for (int i = 0; i < 3; ++i) s[o++] = e.command[i];
    for (int i = 0; i < 3; ++i) s[o++] = e.command_filtered[i];
    for (int i = 0; i < NJ; ++i) s[o++] = e.jointRef_[i];
    for (int i = 0; i < NJ; ++i) s[o++] = e.jointDotRef_[i];
    for (int i = 0; i < NJ; ++i) s[o++] = e.EndEffectorRef_[i];
    s[o++] = e.t0_; s[o++] = e.frame_idx;
    for (int i = 0; i < 4; ++i) s[o++] = e.contact_filtered[i];
    for (int i = 0; i < 35; ++i) s[o++] = e.obDouble_[i];
    for (int i = 0; i < 35; ++i) s[o++] = e.obDouble_last_[i];
    for (int i = 0; i < NJ; ++i) s[o++] = e.gf_[6 + i];
    s[o++] = (double)e.itera;
}
template <typename T> void set_state(Env<T>& e, const double* s) {
    int o = 0;
    for (int i = 0; i < NQ; ++i) e.gc_[i] = T(s[o++]);
    for (int i = 0; i < NV; ++i) e.gv_[i] = T(s[o++]);
    for (int i = 0; i < NJ; ++i) e.pTarget12Last_[i] = T(s[o++]);
    for (int i = 0; i < NJ; ++i) { e.torque_last[i] = T(s[o++]); e.torque[i] = e.torque_last[i]; }   // between steps torque == torque_last (ENV:1511-1515); matters only when a step runs zero substeps
    for (int i = 0; i < 3; ++i) e.command[i] = T(s[o++]);
    for (int i = 0; i < 3; ++i) e.command_filtered[i] = T(s[o++]);
    for (int i = 0; i < NJ; ++i) { e.jointRef_[i] = T(s[o++]); e.jointRefLast_[i] = e.jointRef_[i]; }
    for (int i = 0; i < NJ; ++i) e.jointDotRef_[i] = T(s[o++]);
    for (int i = 0; i < NJ; ++i) e.EndEffectorRef_[i] = T(s[o++]);
    e.t0_ = T(s[o++]); e.frame_idx = (int)s[o++];
    for (int i = 0; i < 4; ++i) { e.contact_filtered[i] = T(s[o++]); e.contact_[i] = e.contact_filtered[i]; }
    for (int i = 0; i < 35; ++i) e.obDouble_[i] = T(s[o++]);
    for (int i = 0; i < 35; ++i) e.obDouble_last_[i] = T(s[o++]);
    for (int i = 0; i < NJ; ++i) e.gf_[6 + i] = T(s[o++]);
    e.itera = (unsigned long)s[o++];
}
template <typename T> void mass_and_h(Env<T>& e, double* M, double* h) {
    typename Env<T>::Kin k; e.kinematics(e.gc_, e.gv_, k);
    T Mm[NV][NV], hh[NV]; e.mass_matrix(k, Mm); e.nonlinearities(k, e.gv_, hh);
    if (M) for (int a = 0; a < NV; ++a) for (int b = 0; b < NV; ++b) M[a * NV + b] = Mm[a][b];
    if (h) for (int a = 0; a < NV; ++a) h[a] = hh[a];
}
template <typename T> void body_kin(Env<T>& e, double* out) {
    // per body: R(9) pc_abs(3) vc(3) w(3) mass(1) Ibody(9)  = 28 doubles
    typename Env<T>::Kin k; e.kinematics(e.gc_, e.gv_, k);
    for (int i = 0; i < NB; ++i) {
        double* o = out + 28 * i;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[3 * r + c] = k.R[i].m[r][c];
        for (int a = 0; a < 3; ++a) { o[9 + a] = k.pc[i][a] + e.gc_[a]; o[12 + a] = k.vc[i][a]; o[15 + a] = k.w[i][a]; }
        o[18] = e.model.b[i].mass;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[19 + 3 * r + c] = e.model.b[i].I.m[r][c];
    }
}
template <typename T> void toe_kin(Env<T>& e, double* out) {  // per foot: pos_abs(3) vel(3)
    typename Env<T>::Kin k; e.kinematics(e.gc_, e.gv_, k);
    for (int l = 0; l < 4; ++l) for (int a = 0; a < 3; ++a) { out[6 * l + a] = k.toe[l][a] + e.gc_[a]; out[6 * l + 3 + a] = k.vtoe[l][a]; }
}
template <typename T> void gait(Env<T>& e, double t, const double* cmd, int is_first, double* jref, double* jdref, double* eeref) {
    e.t0_ = T(t); e.frame_idx = 0;
    for (int i = 0; i < 3; ++i) e.command_filtered[i] = T(cmd[i]);
    e.gait_generator_manual(is_first != 0);
    for (int j = 0; j < NJ; ++j) { jref[j] = e.jointRef_[j]; jdref[j] = e.jointDotRef_[j]; eeref[j] = e.EndEffectorRef_[j]; }
}
template <typename T> void contact_info(Env<T>& e, double* out) {
    // [0..3] foot_in_contact, [4..15] foot_impulse, [16] n_contacts, [17] sweeps, [18..21] force_norm, [22..25] vel_norm
    for (int l = 0; l < 4; ++l) { out[l] = e.foot_in_contact[l]; for (int a = 0; a < 3; ++a) out[4 + 3 * l + a] = e.foot_impulse[l][a]; }
    out[16] = e.n_contacts; out[17] = e.last_solver_sweeps;
    for (int l = 0; l < 4; ++l) { out[18 + l] = e.contact_force_norm[l]; out[22 + l] = e.contact_vel_norm[l]; }
}
template <typename T> void reward_terms(Env<T>& e, double* out) {
    out[0] = e.EndEffectorReward; out[1] = e.BodyCenterReward; out[2] = e.BodyAttitudeReward; out[3] = e.JointReward;
    out[4] = e.JointDotReward; out[5] = e.VelocityReward; out[6] = e.TorqueReward; out[7] = e.ContactReward;
}
template <typename T> void model_params(Env<T>& e, double* out) {
    // mu, restitution, threshold, then per body: mass, com(3), off(3)  -> 3 + 13*7
    out[0] = e.mu; out[1] = e.restitution; out[2] = e.rest_threshold;
    for (int i = 0; i < NB; ++i) { double* o = out + 3 + 7 * i; o[0] = e.model.b[i].mass; for (int a = 0; a < 3; ++a) { o[1 + a] = e.model.b[i].com[a]; o[4 + a] = e.model.b[i].off[a]; } }
}
}  // namespace

#define H(h) (reinterpret_cast<Handle*>(h))
#define DISPATCH(h, expr_d, expr_f) do { if (H(h)->precision == 0) { auto& V = H(h)->d; (void)V; expr_d; } else { auto& V = H(h)->f; (void)V; expr_f; } } while (0)

// meteor sphere (Crutial: True): [p(3), v(3), mode, radius, mass]
template <typename T> void get_meteor(const Env<T>& e, double* s) { for (int i = 0; i < 3; ++i) { s[i] = e.met_p[i]; s[3 + i] = e.met_v[i]; } s[6] = e.met_mode; s[7] = e.met_r; s[8] = e.met_m; }
template <typename T> void set_meteor(Env<T>& e, const double* s) { for (int i = 0; i < 3; ++i) { e.met_p[i] = T(s[i]); e.met_v[i] = T(s[3 + i]); } e.met_mode = int(s[6]); e.met_r = T(s[7]); e.met_m = T(s[8]); }
extern "C" {

void* bp5o_create(const char* cfg, int precision, int env_offset) {
    Handle* h = new Handle(); h->precision = precision;
    try {
        Cfg c = Cfg::parse(cfg);
        if (precision == 0) h->d.create(c, env_offset); else h->f.create(c, env_offset);
    } catch (const std::exception& e) { fprintf(stderr, "bp5o_create: %s\n", e.what()); delete h; return nullptr; }
    return h;
}
// same as bp5o_create with a robot description other than the shipped one (40 numbers, see Model::set_compact)
void* bp5o_create_with_model(const char* cfg, int precision, int env_offset, const double* model40) {
    Handle* h = new Handle(); h->precision = precision;
    try {
        Cfg c = Cfg::parse(cfg);
        if (precision == 0) h->d.create(c, env_offset, model40); else h->f.create(c, env_offset, model40);
    } catch (const std::exception& e) { fprintf(stderr, "bp5o_create_with_model: %s\n", e.what()); delete h; return nullptr; }
    return h;
}
void bp5o_destroy(void* h) { delete H(h); }
int bp5o_state_dim() { return STATE_DIM; }
int bp5o_num_envs(void* h) { return H(h)->precision == 0 ? (int)H(h)->d.envs.size() : (int)H(h)->f.envs.size(); }
void bp5o_set_tick(void* h, unsigned tick) { DISPATCH(h, V.tick = tick, V.tick = tick); }
unsigned bp5o_get_tick(void* h) { return H(h)->precision == 0 ? H(h)->d.tick : H(h)->f.tick; }
void bp5o_set_ref(void* h, const float* ref, int rows) { DISPATCH(h, V.set_ref(ref, rows), V.set_ref(ref, rows)); }
void bp5o_set_terrain(void* h, const float* hf, int nx, int ny, double xs, double ys, double cx, double cy) { DISPATCH(h, { V.set_terrain(hf, nx, ny, xs, ys, cx, cy); if (V.tick == 0) V.reset_all(); }, { V.set_terrain(hf, nx, ny, xs, ys, cx, cy); if (V.tick == 0) V.reset_all(); }); }
void bp5o_reset(void* h, float* ob) { DISPATCH(h, { V.reset_all(); V.observe(ob); }, { V.reset_all(); V.observe(ob); }); }
void bp5o_observe(void* h, float* ob) { DISPATCH(h, V.observe(ob), V.observe(ob)); }
int bp5o_step(void* h, const float* action, float* ob, float* reward, uint8_t* done, float* extra) {
    try { DISPATCH(h, V.step(action, ob, reward, done, extra), V.step(action, ob, reward, done, extra)); }
    catch (const std::exception& e) { fprintf(stderr, "bp5o_step: %s\n", e.what()); return -1; }
    return 0;
}
void bp5o_get_meteor(void* h, int env, double* s) { DISPATCH(h, get_meteor(V.envs[env], s), get_meteor(V.envs[env], s)); }
void bp5o_set_meteor(void* h, int env, const double* s) { DISPATCH(h, set_meteor(V.envs[env], s), set_meteor(V.envs[env], s)); }
void bp5o_get_state(void* h, int env, double* s) { DISPATCH(h, get_state(V.envs[env], s), get_state(V.envs[env], s)); }
void bp5o_set_state(void* h, int env, const double* s) { DISPATCH(h, set_state(V.envs[env], s), set_state(V.envs[env], s)); }
void bp5o_mass_and_h(void* h, int env, double* M, double* hh) { DISPATCH(h, mass_and_h(V.envs[env], M, hh), mass_and_h(V.envs[env], M, hh)); }
void bp5o_body_kin(void* h, int env, double* out) { DISPATCH(h, body_kin(V.envs[env], out), body_kin(V.envs[env], out)); }
void bp5o_toe_kin(void* h, int env, double* out) { DISPATCH(h, toe_kin(V.envs[env], out), toe_kin(V.envs[env], out)); }
void bp5o_gait(void* h, int env, double t, const double* cmd, int is_first, double* jref, double* jdref, double* eeref) {
    DISPATCH(h, gait(V.envs[env], t, cmd, is_first, jref, jdref, eeref), gait(V.envs[env], t, cmd, is_first, jref, jdref, eeref));
}
int bp5o_integrate(void* h, int env, const double* tau12) {
    try {
        if (H(h)->precision == 0) { H(h)->d.envs[env].integrate(tau12); }
        else { float t[12]; for (int i = 0; i < 12; ++i) t[i] = (float)tau12[i]; H(h)->f.envs[env].integrate(t); }
    } catch (const std::exception& e) { fprintf(stderr, "bp5o_integrate: %s\n", e.what()); return -1; }
    return 0;
}
void bp5o_contact_info(void* h, int env, double* out) { DISPATCH(h, contact_info(V.envs[env], out), contact_info(V.envs[env], out)); }
// [geo, rest, term, cone] decision margins of the last control step (see Env::margin_*)
void bp5o_margins(void* h, int env, double* out) { DISPATCH(h, { out[0] = V.envs[env].margin_geo; out[1] = V.envs[env].margin_rest; out[2] = V.envs[env].margin_term; out[3] = V.envs[env].margin_cone; }, { out[0] = V.envs[env].margin_geo; out[1] = V.envs[env].margin_rest; out[2] = V.envs[env].margin_term; out[3] = V.envs[env].margin_cone; }); }
void bp5o_reward_terms(void* h, int env, double* out) { DISPATCH(h, reward_terms(V.envs[env], out), reward_terms(V.envs[env], out)); }
void bp5o_model_params(void* h, int env, double* out) { DISPATCH(h, model_params(V.envs[env], out), model_params(V.envs[env], out)); }
int bp5o_is_terminal(void* h, int env) {
    float tr; return H(h)->precision == 0 ? (int)H(h)->d.envs[env].isTerminalState(tr) : (int)H(h)->f.envs[env].isTerminalState(tr);
}
// the pure helper functions, for the comparison with the compiled slices of the reference (oracle/_ref, tests/test_oracle_ref_slices.py)
// which: 0 sampling_reshape(a)  1 gauss(a,b,c)  2 smooth_function(a,b,c)  3 smooth_function2(a,b,c)
double bp5o_helper(int which, double a, double b, double c) {
    switch (which) { case 0: return sampling_reshape(a); case 1: return gauss<double>(a, b, c); case 2: return smooth_function<double>(a, b, c); default: return smooth_function2<double>(a, b, c); }
}
// ENV:1687-1751 on env 0's leg lengths / max_len; theta[3] in/out
void bp5o_ik(void* h, double x, double y, double z, int is_right, double* theta) {
    if (H(h)->precision == 0) { auto& e = H(h)->d.envs[0]; e.inverse_kinematics(x, y, z, e.l_hip_, e.l_thigh_, e.l_calf_, theta, is_right != 0); }
    else { auto& e = H(h)->f.envs[0]; float th[3] = {(float)theta[0], (float)theta[1], (float)theta[2]}; e.inverse_kinematics((float)x, (float)y, (float)z, e.l_hip_, e.l_thigh_, e.l_calf_, th, is_right != 0); for (int i = 0; i < 3; ++i) theta[i] = th[i]; }
}
// ENV:1273-1305 on env 0's motor constants: torque[12] in/out, gv[18]
void bp5o_torque_clamp(void* h, double* torque, const double* gv) {
    if (H(h)->precision == 0) { auto& e = H(h)->d.envs[0]; for (int i = 0; i < 12; ++i) e.torque[i] = torque[i]; e.torque_clamp(gv); for (int i = 0; i < 12; ++i) torque[i] = e.torque[i]; }
    else { auto& e = H(h)->f.envs[0]; float g[18]; for (int i = 0; i < 18; ++i) g[i] = (float)gv[i]; for (int i = 0; i < 12; ++i) e.torque[i] = (float)torque[i]; e.torque_clamp(g); for (int i = 0; i < 12; ++i) torque[i] = e.torque[i]; }
}
// CPU arm of bench.py: the two-tower 2x48 LSTM act (CustomerLstmNN.py:112-156 + run_bp_v5.py:117-185; float32 like the
// reference's TF graph) over all envs with OpenMP, one env per iteration.  params = the 19 arrays of the checkpoint flattened in
// their order (70 741 floats); state [n,384] updated in place; eps [n,12] standard normal draws (NULL: deterministic).
static inline float sigm(float x) { return 1.0f / (1.0f + std::exp(-x)); }
static void lstm_cell_f(const float* x, int nin, float* c, float* h, const float* wx, const float* wh, const float* b) {
    float g[192];
    for (int j = 0; j < 192; ++j) g[j] = b[j];
    for (int k = 0; k < nin; ++k) { const float xv = x[k]; const float* w = wx + (size_t)k * 192; for (int j = 0; j < 192; ++j) g[j] += xv * w[j]; }
    for (int k = 0; k < 48; ++k) { const float hv = h[k]; const float* w = wh + (size_t)k * 192; for (int j = 0; j < 192; ++j) g[j] += hv * w[j]; }
    for (int u = 0; u < 48; ++u) {          // gate order i, f, o, g (CustomerLstmNN.py:119-122)
        const float i = sigm(g[u]), f = sigm(g[48 + u]), o = sigm(g[96 + u]), gg = std::tanh(g[144 + u]);
        c[u] = f * c[u] + i * gg; h[u] = o * std::tanh(c[u]);
    }
}
void bp5o_lstm_act(const float* params, int n, const float* obs, float* state, const uint8_t* done, const float* eps, float* action, float* value, float* neglogp, int threads) {
    const float* p = params; const int in[4] = {35, 48, 35, 48};
    const float *wx[4], *wh[4], *bb[4];
    for (int i = 0; i < 4; ++i) { wx[i] = p; p += in[i] * 192; wh[i] = p; p += 48 * 192; bb[i] = p; p += 192; }
    const float* vf_w = p; p += 48; const float* vf_b = p; p += 1; const float* pi_w = p; p += 48 * 12; const float* pi_b = p; p += 12; const float* logstd = p;
    #pragma omp parallel for schedule(static) num_threads(threads)
    for (int e = 0; e < n; ++e) {
        float* s = state + (size_t)e * 384; const float* ob = obs + (size_t)e * 35;
        if (done && done[e]) for (int k = 0; k < 384; ++k) s[k] = 0.f;                     // SB lstm(): c *= 1 - m, h *= 1 - m
        for (int tower = 0; tower < 2; ++tower) {                                          // state = [c0,h0,c1,h1]_pi || [c0,h0,c1,h1]_V
            float* t = s + 192 * tower;
            lstm_cell_f(ob, 35, t, t + 48, wx[2 * tower], wh[2 * tower], bb[2 * tower]);
            lstm_cell_f(t + 48, 48, t + 96, t + 144, wx[2 * tower + 1], wh[2 * tower + 1], bb[2 * tower + 1]);
        }
        const float* hp = s + 144; const float* hv = s + 192 + 144;
        float v = vf_b[0]; for (int k = 0; k < 48; ++k) v += hv[k] * vf_w[k];
        value[e] = v;
        float nl = 0.5f * 12.f * std::log(2.0f * 3.14159265358979f);
        for (int j = 0; j < 12; ++j) {
            float m = pi_b[j]; for (int k = 0; k < 48; ++k) m += hp[k] * pi_w[k * 12 + j];
            const float sd = std::exp(logstd[j]); const float a = eps ? m + sd * eps[(size_t)e * 12 + j] : m;
            action[(size_t)e * 12 + j] = a; const float zz = (a - m) / sd; nl += 0.5f * zz * zz + logstd[j];
        }
        neglogp[e] = nl;
    }
}

// Philox block exposed so tests can pin the RNG against the published Random123 known-answer vectors
void bp5o_philox(unsigned seed, unsigned env, unsigned tick, unsigned purpose, unsigned* out) { Philox::gen(seed, env, tick, purpose, out); }
// raw Philox4x32-10 with arbitrary counter/key, for the Random123 KAT
void bp5o_philox_raw(const unsigned* ctr, const unsigned* key, unsigned* out) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1; Philox::mulhilo(0xD2511F53u, c0, hi0, lo0); Philox::mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0; c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// CPU baseline timing: `steps` control steps over all envs with the same OpenMP
// structure as VEC:273; actions = clip(N(0, sigma^2), +-1) (sigma = 0 -> zeros).  Returns seconds spent
// inside step() only (std::chrono::steady_clock), env-steps = steps * num_envs.
// F_step measured by the oracle's op counter.  Builds the same vec-env on the counting scalar (single thread), runs `warm`
// control steps, then counts every floating-point operation of `steps` further steps (actions ~ clip(N(0, sigma^2))).
// out[0..5] = add/sub, mul, div, sqrt, transcendental, compare -- each per env-step; returns add + mul + div + sqrt + transcendental.
double bp5o_count_flops(const char* cfg, int warm, int steps, double sigma, unsigned seed, double* out6) {
    try {
        Cfg c = Cfg::parse(cfg);
        VecEnv<Counted> V; V.create(c, 0);
        int n = (int)V.envs.size();
        std::vector<float> act((size_t)n * 12, 0.f), ob((size_t)n * 35), rew(n), extra((size_t)n * 6);
        std::vector<uint8_t> done(n);
        std::mt19937 gen(seed); std::normal_distribution<float> nd(0.f, 1.f);
        V.reset_all();
        for (int s = 0; s < warm + steps; ++s) {
            if (s == warm) op_count() = OpCount();
            if (sigma > 0) for (auto& a : act) { float v = (float)sigma * nd(gen); a = v > 1.f ? 1.f : (v < -1.f ? -1.f : v); }
            V.step(act.data(), ob.data(), rew.data(), done.data(), extra.data());
        }
        const OpCount& k = op_count(); const double d = (double)n * steps;
        if (out6) { out6[0] = k.add / d; out6[1] = k.mul / d; out6[2] = k.div / d; out6[3] = k.sqrt / d; out6[4] = k.trans / d; out6[5] = k.cmp / d; }
        return (k.add + k.mul + k.div + k.sqrt + k.trans) / d;
    } catch (const std::exception& e) { fprintf(stderr, "bp5o_count_flops: %s\n", e.what()); return -1.0; }
}

double bp5o_time_steps(void* h, int steps, double sigma, unsigned seed, long* n_done_out) {
    int n = bp5o_num_envs(h);
    std::vector<float> act((size_t)n * 12, 0.f), ob((size_t)n * 35), rew(n), extra((size_t)n * 6);
    std::vector<uint8_t> done(n);
    std::mt19937 gen(seed); std::normal_distribution<float> nd(0.f, 1.f);
    double total = 0; long ndone = 0;
    for (int s = 0; s < steps; ++s) {
        if (sigma > 0) for (auto& a : act) { float v = (float)sigma * nd(gen); a = v > 1.f ? 1.f : (v < -1.f ? -1.f : v); }
        auto t0 = std::chrono::steady_clock::now();
        bp5o_step(h, act.data(), ob.data(), rew.data(), done.data(), extra.data());
        auto t1 = std::chrono::steady_clock::now();
        total += std::chrono::duration<double>(t1 - t0).count();
        for (int i = 0; i < n; ++i) ndone += done[i];
    }
    if (n_done_out) *n_done_out = ndone;
    return total;
}

}  // extern "C"
