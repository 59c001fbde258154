// Host shim: the C ABI of include/irrl_b200.h over the sm_100a kernels.
// This is the thin replacement of VectorizedEnvironment<ENVIRONMENT> (VEC:127-382) + the pybind module
// (raisim_gym.cpp:14-47): it owns the device state, parses the YAML `environment:` map (ENV:1594-1659,
// VEC:136-171), launches kernels on one stream and stages host buffers through pinned memory.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/irrl_b200.h"
#include "irrl_diag.h"
#include "env_kernels.h"
#include "urdf_reader.h"

using namespace irrl;

namespace {
// NVTX range for the lifetime of a host entry point (shows up in Nsight Systems / ncu --nvtx; no cost without a profiler attached)
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(-2, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

// ------------------------------------------------------------------ minimal YAML flat-map reader
// Accepts what run_bp_v5.py:205-207 produces (a block-style mapping of scalars), optionally nested under an
// `environment:` key when a whole config file is passed.
struct YamlMap {
    std::map<std::string, std::string> kv;
    static std::string trim(const std::string& s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? "" : s.substr(a, b - a + 1);
    }
    bool parse(const std::string& text, std::string& err) {
        std::istringstream in(text); std::string line; bool has_env = false; std::map<std::string, std::string> top, env;
        bool in_env = false;
        while (std::getline(in, line)) {
            size_t hash = std::string::npos; bool q = false;
            for (size_t i = 0; i < line.size(); ++i) { if (line[i] == '\'' || line[i] == '"') q = !q; if (line[i] == '#' && !q && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) { hash = i; break; } }
            if (hash != std::string::npos) line = line.substr(0, hash);
            if (trim(line).empty() || trim(line) == "---") continue;
            size_t indent = line.find_first_not_of(" \t");
            size_t colon = line.find(':');
            if (colon == std::string::npos) { err = "YAML: cannot parse line '" + line + "'"; return false; }
            std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
            if (key.size() >= 2 && (key.front() == '"' || key.front() == '\'')) key = key.substr(1, key.size() - 2);
            if (val.size() >= 2 && (val.front() == '"' || val.front() == '\'') && val.back() == val.front()) val = val.substr(1, val.size() - 2);
            if (indent == 0) { in_env = false; if (key == "environment" && val.empty()) { in_env = true; has_env = true; continue; } top[key] = val; }
            else { if (in_env) env[key] = val; else top[key] = val; }
        }
        kv = has_env ? env : top;
        return true;
    }
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    static bool to_bool(const std::string& v, bool& out) {
        std::string s; for (char c : v) s += (char)tolower(c);
        if (s == "true" || s == "yes" || s == "on" || s == "1") { out = true; return true; }
        if (s == "false" || s == "no" || s == "off" || s == "0") { out = false; return true; }
        return false;
    }
};

struct irrl_env_impl {
    EnvParams P{};
    DevState S{};
    int device = 0;
    cudaStream_t stream = nullptr; bool own_stream = false;
    uint32_t tick = 0;
    bool initialised = false;
    std::string resource_dir, ref_path;
    double max_time_d = 0, control_dt_d = 0, sim_dt_d = 0, period_d = 0;
    // heightfield (host copy + generator settings from the YAML)
    std::vector<float> terrain_h; std::string terrain_kind = "perlin"; double stair_rise = 0.08, stair_run = 0.3, stair_start = 1.0; int terrain_seed = 0; float* d_terrain = nullptr;   // YAML values in double: integer counts are derived like the reference does
    std::vector<std::string> extra_names;
    // staging (device + pinned host)
    float *d_action = nullptr, *d_ob = nullptr, *d_reward = nullptr, *d_extra = nullptr, *d_ep_ret = nullptr, *d_scratch = nullptr;
    uint8_t* d_done = nullptr; int* d_ep_len = nullptr;
    float* d_ref = nullptr;
    unsigned char* h_pin = nullptr; size_t h_pin_bytes = 0; size_t scratch_floats = 0;
    std::vector<void*> allocs;
    // irrl_act_step: device copies of the act outputs [action | clipped | value | neglogp], chunk streams and their join events
    float* d_fused_act = nullptr; cudaStream_t chunk_stream[8] = {}; cudaEvent_t chunk_done[8] = {}; cudaEvent_t fork_ev = nullptr; int n_chunk_streams = 0;
    cudaStream_t side_stream[8] = {}; cudaEvent_t act_done[8] = {}, side_done[8] = {};      // act results go home behind the step kernel of their chunk
    // optional per-kernel timing of irrl_rollout (CUDA events on the env's stream)
    bool profiling = false; std::vector<cudaEvent_t> prof_events; size_t prof_cursor = 0; double prof_act_ms = 0, prof_step_ms = 0; long prof_count = 0;
};
struct irrl_policy_impl {
    int device = 0;
    float* d_params = nullptr;
    float* d_derived = nullptr;   // 4 x ([96][192] + [192]) gate-interleaved copies for the act kernel
    unsigned char* d_tcblob = nullptr;   // 2 x TC_BLOB_BYTES hi/lo-split UMMA-layout weights for the tcgen05 act kernel
    PolicyWeights W{};
    // staging for host callers
    int cap = 0; float *d_obs = nullptr, *d_state = nullptr, *d_action = nullptr, *d_clipped = nullptr, *d_value = nullptr, *d_nlp = nullptr; uint8_t* d_done = nullptr;
    unsigned char* h_pin = nullptr;
};

// Registry of the ranges page-locked through irrl_host_register.  With IRRL_ZERO_COPY=1 (outputs and inputs) or 2 (inputs only)
// the kernels use those buffers in place (mapped pinned memory) instead of cudaMemcpy staging.  Measured on B200 / PCIe 5:
// in-place outputs are 35 % SLOWER end to end (the step kernel's scalar observation stores become small PCIe writes) and
// in-place inputs are neutral, so the default is 0 = DMA copies; the switch stays for experiments.
struct PinnedRange { size_t bytes; char* dev; };
static std::map<uintptr_t, PinnedRange> g_pinned;
static std::mutex g_pinned_mu;
static int zero_copy_mode() { static int m = [] { const char* e = getenv("IRRL_ZERO_COPY"); return e ? atoi(e) : 0; }(); return m; }
// device-visible alias of [p, p+bytes) if the range lies inside one registered block, else nullptr
template <typename T> static T* mapped_alias(T* p, size_t bytes) {
    if (!p || zero_copy_mode() == 0) return nullptr;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    auto it = g_pinned.upper_bound(a);
    if (it == g_pinned.begin()) return nullptr;
    --it;
    if (a + bytes > it->first + it->second.bytes) return nullptr;
    return reinterpret_cast<T*>(it->second.dev + (a - it->first));
}
// [p, p+bytes) lies inside a block this library page-locked itself: no driver query needed to classify it
static bool in_registry(const void* p, size_t bytes) {
    if (!p) return false;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    auto it = g_pinned.upper_bound(a);
    if (it == g_pinned.begin()) return false;
    --it;
    return a + bytes <= it->first + it->second.bytes;
}

bool is_device_ptr(const void* p) {
    if (!p || in_registry(p, 1)) return false;
    cudaPointerAttributes a; cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// page-locked host memory (cudaHostAlloc / cudaHostRegister): DMA can target it directly, no staging copy needed
bool is_pinned_host_ptr(const void* p, size_t bytes = 1) {
    if (!p) return false;
    if (in_registry(p, bytes ? bytes : 1)) return true;
    // both ends of the range must be page-locked (a stale registration may cover only part of a recycled allocation)
    const void* ends[2] = {p, static_cast<const char*>(p) + (bytes ? bytes - 1 : 0)};
    for (const void* q : ends) {
        cudaPointerAttributes a; cudaError_t e = cudaPointerGetAttributes(&a, q);
        if (e != cudaSuccess) { cudaGetLastError(); return false; }
        if (a.type != cudaMemoryTypeHost) return false;
    }
    return true;
}

static cudaError_t wait_stream(cudaStream_t st) { return cudaStreamSynchronize(st); }

template <typename T> int dev_alloc(irrl_env_impl* E, T** p, size_t n) {
    void* q = nullptr; CUDA_OK(cudaMalloc(&q, n * sizeof(T))); CUDA_OK(cudaMemset(q, 0, n * sizeof(T)));
    E->allocs.push_back(q); *p = reinterpret_cast<T*>(q); return 0;
}

int read_cfg(irrl_env_impl* E, const YamlMap& y) {
    EnvParams& P = E->P;
    std::string missing;
    auto num = [&](const char* k, double& out) { auto it = y.kv.find(k); if (it == y.kv.end()) { missing = k; return false; }
        char* end = nullptr; out = strtod(it->second.c_str(), &end); if (end == it->second.c_str()) { bool b; if (YamlMap::to_bool(it->second, b)) { out = b; return true; } missing = std::string(k) + " (not a number: '" + it->second + "')"; return false; } return true; };
    auto flg = [&](const char* k, int& out) { auto it = y.kv.find(k); if (it == y.kv.end()) { missing = k; return false; }
        bool b; if (!YamlMap::to_bool(it->second, b)) { missing = std::string(k) + " (not a bool: '" + it->second + "')"; return false; } out = b ? 1 : 0; return true; };
    double d; int f;
#define NUM(key, field) if (!num(key, d)) goto bad; field = (float)d;
#define FLG(key, field) if (!flg(key, f)) goto bad; field = f;
#define IGN(key) if (!y.has(key)) { missing = key; goto bad; }
    {
        // VEC:146-171
        if (!num("num_envs", d)) goto bad; P.N = (int)d;
        IGN("num_threads");
        double sim_dt, ctl_dt; if (!num("simulation_dt", sim_dt)) goto bad; if (!num("control_dt", ctl_dt)) goto bad;
        P.sim_dt = (float)sim_dt; P.control_dt = (float)ctl_dt; P.loop_count = int(ctl_dt / sim_dt + 1e-10);   // ENV:711
        E->control_dt_d = ctl_dt; E->sim_dt_d = sim_dt; P.control_dt_d = ctl_dt;
        if (!num("seedd", d)) goto bad; P.seed = (uint32_t)(int)d;                                               // VEC:171
        // ENV:1598-1613
        NUM("abad", P.abad) NUM("period", P.period) E->period_d = d; P.period_d = d; P.inv_period_d = 1.0 / d; P.disturb_every = int(d / ctl_dt * 10.0);   /* ENV:746, double like the reference */ P.meteor_every = int(5.0 * d / ctl_dt);   /* ENV:731 */ NUM("lam", P.lam) NUM("stand_height", P.stand_height) NUM("up_height", P.up_height_max)
        IGN("down_height") IGN("gait_step")
        NUM("Vx", P.Vx_max) P.Vx_min = 0.f;                                                                      // ENV:1606-1607, 2054
        NUM("Vy", P.Vy_max) P.Vy_min = -P.Vy_max; NUM("Omega", P.omega_max) P.omega_min = -P.omega_max;
        NUM("LeanFront", P.lean_front) NUM("LeanHind", P.lean_hind)
        // ENV:1616-1629
        FLG("Terrain", P.flag_terrain) FLG("Manual", P.flag_manual) FLG("Crutial", P.flag_crucial)
        FLG("Filter", P.flag_filter) IGN("Camera") FLG("StochasticDynamics", P.flag_stochastic) FLG("HeightVariable", P.flag_height_variable)
        FLG("TimeBasedContact", P.flag_time_contact) FLG("ManualTraj", P.flag_manual_traj) IGN("MotorDynamics") FLG("ObsFilter", P.flag_obs_filter)
        FLG("WILDCAT", P.flag_wildcat) FLG("ForceDisturbance", P.flag_force_dist) IGN("Convert2Torque")
        // ENV:1632-1639
        NUM("terminalRewardCoeff", P.terminal_coeff) NUM("EndEffectorRewardCoeff", P.ee_coeff) NUM("BodyPosRewardCoeff", P.pos_coeff)
        NUM("BodyAttitudeRewardCoeff", P.atti_coeff) NUM("JointRewardCoeff", P.joint_coeff) NUM("VelRewardCoeff", P.vel_coeff)
        NUM("TorqueCoeff", P.torque_coeff) NUM("ContactCoeff", P.contact_coeff)
        // ENV:1643-1658
        NUM("Stiffness", P.stiffness) IGN("Stiffness_Low") NUM("AbadRatio", P.abad_ratio) NUM("Damping", P.damping)
        double freq; if (!num("Freq", freq)) goto bad;
        NUM("max_time", P.max_time) E->max_time_d = d; if (!num("CubeNum", d)) goto bad; P.num_cube = (int)d; IGN("FPS") NUM("ActionNoise", P.action_noise) NUM("ObsNoise", P.noise_flag)
        if (!num("GaitType", d)) goto bad; P.gait_type = (int)d;
        NUM("MotorMaxTorque", P.motor_max_torque) NUM("MotorCriticalSpeed", P.motor_crit_speed) NUM("MotorMaxSpeed", P.motor_max_speed)
        P.filter_para = P.flag_filter ? (float)(1.0 - freq * ctl_dt) : 0.f;                                      // ENV:396
        P.obs_filter_alpha = P.flag_obs_filter ? (float)(2.0 * 3.14 * ctl_dt * 20.0 / (2.0 * 3.14 * ctl_dt * 20.0 + 1.0)) : 1.f;   // ENV:425-426, 2026
        if (y.has("RefTraj")) E->ref_path = y.kv.at("RefTraj");
        if (y.has("terrain_kind")) E->terrain_kind = y.kv.at("terrain_kind");
        // optional solver / model switches of this implementation
        auto opt = [&](const char* k, double dflt) { double v; return y.has(k) && num(k, v) ? v : dflt; };
        P.joint_damping = (float)opt("joint_damping", 0.01);                                                     // URDF:56
        P.solver_iters = (int)opt("solver_iters", 30); P.jacobi_sweeps = (int)opt("jacobi_sweeps", 10); P.slide_iters = (int)opt("slide_iters", 1); P.solver_tol = (float)opt("solver_tol", 1e-5);
        E->stair_rise = opt("stair_rise", 0.08); E->stair_run = opt("stair_run", 0.3); E->stair_start = opt("stair_start", 1.0); E->terrain_seed = (int)opt("terrain_seed", 0);
        P.mu = (float)opt("friction", 0.6); P.restitution = (float)opt("restitution", 0.2); P.rest_threshold = (float)opt("restitution_threshold", 0.01);   // ENV:433
    }
    if (P.N <= 0) return fail(-3, "num_envs must be positive");
    if (P.slide_iters != 1) return fail(-3, "slide_iters: the CUDA contact solve takes exactly one fixed-point step on the sliding direction (other values exist in the CPU oracle only)");
    // ForceDisturbance: with Manual the step kernel perturbs the base state every 10 gait periods (state_disturbance, ENV:912-940);
    // without Manual force_attack(random() < 0.0027) never fires -> zero external force.
    switch (P.gait_type) {                                                                                       // ENV:398-409
        case 0: P.phase[0] = 0.5f; P.phase[1] = 0.f; P.phase[2] = 0.f; P.phase[3] = 0.5f; break;
        case 1: P.phase[0] = 0.5f; P.phase[1] = 0.5f; P.phase[2] = 0.f; P.phase[3] = 0.f; break;
        case 2: P.phase[0] = 0.f; P.phase[1] = 0.25f; P.phase[2] = 0.5f; P.phase[3] = 0.75f; break;
        default: P.phase[0] = P.phase[1] = P.phase[2] = P.phase[3] = 0.f;
    }
    return 0;
bad:
    return fail(-3, "Node cfg[\"" + missing + "\"] doesn't exist");   // GYM:41-42 READ_YAML
#undef NUM
#undef FLG
#undef IGN
}

void model_defaults(EnvParams& P) {
    // hard-coded members ENV:1949-1952, 1988-1999, 2043
    P.l_thigh = 0.209f; P.l_calf = 0.2175f; P.l_hip = 0.085f;
    P.max_len = (float)std::sqrt(0.085 * 0.085 + (0.2175 + 0.209) * (0.2175 + 0.209));                           // ENV:395
    P.joint_noise = 0.002f; P.joint_vel_noise = 0.8f; P.posture_sigma = 0.02f; P.omega_sigma = 0.5f; P.cmd_update = 0.995f;
    // URDF
    P.I0[0] = 0.016269f; P.I0[1] = 0.050813f; P.I0[2] = 0.060989f;                                              // URDF:21
    P.I1[0] = 0.000391f; P.I1[1] = 0.000739f; P.I1[2] = 0.000488f;                                              // URDF:64
    P.I2[0] = 0.001724f; P.I2[1] = 0.001907f; P.I2[2] = 0.000468f; P.I2[3] = 0.000228f;                         // URDF:92
    const double ms = 0.064, mt = 0.05, zs = -0.0865, zt = -0.19;                                               // URDF:115-119, 152-162
    const double zc = (ms * zs + mt * zt) / (ms + mt), ds = zs - zc, dt = zt - zc;
    P.I3[0] = (float)(0.000716 + ms * ds * ds + 0.000025 + mt * dt * dt);
    P.I3[1] = (float)(0.000721 + ms * ds * ds + 0.000025 + mt * dt * dt);
    P.I3[2] = (float)(0.000012 + 0.000025);
    P.rotor[0] = 0.003708f; P.rotor[1] = 0.003708f; P.rotor[2] = 0.008966f;                                     // URDF:56,84,110
    P.off1x = 0.212f; P.off1y = 0.051f; P.off2y = 0.085f;                                                       // URDF:52,80
    P.toe_z = -0.19f; P.toe_r = 0.0275f;                                                                        // URDF:162,148
    P.box_half[0] = 0.15f; P.box_half[1] = 0.1f; P.box_half[2] = 0.05f;                                         // URDF:26
    P.gravity = 9.81f;
    P.m0 = 3.72f; P.com0[0] = 0.f; P.com0[1] = 0.f; P.com0[2] = -0.003f;                                        // URDF:18-20
    P.m1 = 0.54f; P.com1[0] = 0.058f; P.com1[1] = 0.00485f; P.com1[2] = 0.f;                                    // URDF:62-63 (x sx, y sy)
    P.m2 = 0.636f; P.com2[0] = 0.f; P.com2[1] = -0.019f; P.com2[2] = -0.01865f;                                 // URDF:90-91 (y sy)
    P.m3 = (float)(ms + mt); P.com3z = (float)zc; P.knee_z = -0.201f;                                           // URDF:106
}

// robot description found under resourceDir (ENV:231): overrides the built-in bp5 constants
void apply_urdf(EnvParams& P, const CompactModel& M) {
    for (int i = 0; i < 3; ++i) { P.I0[i] = (float)M.I0[i]; P.I1[i] = (float)M.I1[i]; P.I3[i] = (float)M.I3[i]; P.rotor[i] = (float)M.rotor[i]; P.box_half[i] = (float)M.box_half[i];
                                  P.com0[i] = (float)M.com0[i]; P.com1[i] = (float)M.com1[i]; P.com2[i] = (float)M.com2[i]; }
    for (int i = 0; i < 4; ++i) P.I2[i] = (float)M.I2[i];
    P.off1x = (float)M.off1x; P.off1y = (float)M.off1y; P.off2y = (float)M.off2y; P.toe_z = (float)M.toe_z; P.toe_r = (float)M.toe_r;
    P.m0 = (float)M.m0; P.m1 = (float)M.m1; P.m2 = (float)M.m2; P.m3 = (float)M.m3; P.com3z = (float)M.com3z; P.knee_z = (float)M.knee_z;
    P.joint_damping = (float)M.joint_damping;
}
// the model fields of EnvParams in a fixed order (irrl_parse_urdf, tests)
void model_to_array(const EnvParams& P, float* o) {
    int k = 0;
    for (int i = 0; i < 3; ++i) o[k++] = P.I0[i]; for (int i = 0; i < 3; ++i) o[k++] = P.I1[i]; for (int i = 0; i < 4; ++i) o[k++] = P.I2[i]; for (int i = 0; i < 3; ++i) o[k++] = P.I3[i];
    for (int i = 0; i < 3; ++i) o[k++] = P.rotor[i];
    o[k++] = P.off1x; o[k++] = P.off1y; o[k++] = P.off2y; o[k++] = P.toe_z; o[k++] = P.toe_r;
    for (int i = 0; i < 3; ++i) o[k++] = P.box_half[i];
    o[k++] = P.joint_damping; o[k++] = P.m0; for (int i = 0; i < 3; ++i) o[k++] = P.com0[i]; o[k++] = P.m1; for (int i = 0; i < 3; ++i) o[k++] = P.com1[i];
    o[k++] = P.m2; for (int i = 0; i < 3; ++i) o[k++] = P.com2[i]; o[k++] = P.m3; o[k++] = P.com3z; o[k++] = P.knee_z;      // 40 values
}

int load_csv(const std::string& path, std::vector<float>& out, int& rows, int& cols) {
    // readCSV_m  VEC:33-76
    std::ifstream in(path); rows = 0; cols = 0;
    if (!in.is_open()) { fprintf(stdout, "Can Not Load Parameter File of %s\n", path.c_str()); return -1; }   // VEC:71-73
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        int c = 0; const char* start = line.c_str();
        for (size_t i = 0; i <= line.size(); ++i) if (i == line.size() || line[i] == ',') { out.push_back((float)atof(start)); start = line.c_str() + i + 1; ++c; }
        if (rows == 0) cols = c; else if (c != cols) return -1;
        ++rows;
    }
    return 0;
}

int ensure_pin(irrl_env_impl* E, size_t bytes) {
    if (bytes <= E->h_pin_bytes) return 0;
    if (E->h_pin) cudaFreeHost(E->h_pin);
    CUDA_OK(cudaMallocHost((void**)&E->h_pin, bytes)); E->h_pin_bytes = bytes; return 0;
}
int ensure_scratch(irrl_env_impl* E, size_t floats) {
    if (floats <= E->scratch_floats) return 0;
    void* q; CUDA_OK(cudaMalloc(&q, floats * sizeof(float))); E->allocs.push_back(q); E->d_scratch = (float*)q; E->scratch_floats = floats; return 0;
}
// copy a device result [count] to the caller (host or device pointer)
int deliver(irrl_env_impl* E, void* user, const void* dev, size_t bytes) {
    if (!user) return 0;
    if (is_device_ptr(user)) { CUDA_OK(cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToDevice, E->stream)); return 0; }
    if (is_pinned_host_ptr(user, bytes)) {
        CUDA_OK(cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, E->stream));
        CUDA_OK(cudaStreamSynchronize(E->stream)); return 0;
    }
    if (int rc = ensure_pin(E, bytes)) return rc;
    CUDA_OK(cudaMemcpyAsync(E->h_pin, dev, bytes, cudaMemcpyDeviceToHost, E->stream));
    CUDA_OK(cudaStreamSynchronize(E->stream));
    memcpy(user, E->h_pin, bytes); return 0;
}
#define ENV(e) irrl_env_impl* E = reinterpret_cast<irrl_env_impl*>(e); if (!E) return fail(-1, "null env"); cudaSetDevice(E->device)
#define NEED_INIT() if (!E->initialised) return fail(-1, "irrl_init() has not been called")

StepArgs make_args(irrl_env_impl* E, const float* action, float* ob, float* reward, uint8_t* done, float* extra) {
    StepArgs a; a.P = E->P; a.S = E->S; a.action = action; a.ob = ob; a.reward = reward; a.done = done; a.extra = extra;
    a.ep_ret_out = E->d_ep_ret; a.ep_len_out = E->d_ep_len; a.tick = E->tick; a.r_begin = 0; a.r_end = E->P.N; return a;
}

// generic probe helper: run `fn` into a device buffer of `per_env` floats and deliver
template <typename F> int probe(irrl_env_impl* E, float* out, int per_env, F fn) {
    size_t n = (size_t)E->P.N * per_env;
    if (is_device_ptr(out)) { fn(out); CUDA_OK(cudaGetLastError()); return 0; }
    if (int rc = ensure_scratch(E, n)) return rc;
    fn(E->d_scratch); CUDA_OK(cudaGetLastError());
    return deliver(E, out, E->d_scratch, n * sizeof(float));
}
}  // namespace

extern "C" {

const char* irrl_last_error(void) { return g_err.c_str(); }
const char* irrl_version(void) { return "irrl_b200 0.1.0 (sm_100a)"; }

int irrl_create(const char* resource_dir, const char* cfg_yaml, int device, int env_offset, irrl_env** out) {
    if (!out || !cfg_yaml) return fail(-1, "null argument");
    *out = nullptr;
    int ndev = 0; cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(-2, "no CUDA device: this library has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(-2, "invalid CUDA device ordinal");
    irrl_env_impl* E = new irrl_env_impl(); E->device = device; E->resource_dir = resource_dir ? resource_dir : "";
    YamlMap y; std::string err;
    if (!y.parse(cfg_yaml, err)) { delete E; return fail(-3, err); }
    model_defaults(E->P);
    if (int rc = read_cfg(E, y)) { delete E; return rc; }
    // ENV:231: the robot description lives in the resource directory.  When it is there it replaces the built-in bp5 constants
    // (a description the compact model cannot represent is an error); when it is not, the built-in constants of that very file are used.
    if (!E->resource_dir.empty()) {
        for (const char* rel : {"/black_panther.urdf", "/urdf/black_panther.urdf"}) {
            const std::string path = E->resource_dir + rel;
            std::ifstream probe(path);
            if (!probe.is_open()) continue;
            CompactModel M; std::string uerr;
            if (int rc = read_urdf(path, M, uerr)) { (void)rc; delete E; return fail(-3, path + ": " + uerr); }
            const float yaml_damping = E->P.joint_damping;
            apply_urdf(E->P, M);
            if (y.has("joint_damping")) E->P.joint_damping = yaml_damping;
            break;
        }
    }
    E->P.env_offset = (uint32_t)env_offset;
    E->extra_names = {"EndEffectorReward(0.15)", "Height_Keep_Reward(0.1)", "base height", "Balance_Keep_Reward(0.1)", "JointReward(0.65)", "VelocityReward(0.2)"};   // ENV:944-949
    *out = reinterpret_cast<irrl_env*>(E);
    return 0;
}

void irrl_destroy(irrl_env* env) {
    irrl_env_impl* E = reinterpret_cast<irrl_env_impl*>(env); if (!E) return;
    cudaSetDevice(E->device);
    if (E->stream) cudaStreamSynchronize(E->stream);
    for (void* p : E->allocs) cudaFree(p);
    if (E->h_pin) cudaFreeHost(E->h_pin);
    for (int i = 0; i < E->n_chunk_streams; ++i) { cudaStreamDestroy(E->chunk_stream[i]); cudaEventDestroy(E->chunk_done[i]); cudaStreamDestroy(E->side_stream[i]); cudaEventDestroy(E->act_done[i]); cudaEventDestroy(E->side_done[i]); }
    if (E->fork_ev) cudaEventDestroy(E->fork_ev);
    if (E->own_stream && E->stream) cudaStreamDestroy(E->stream);
    delete E;
}

int irrl_set_stream(irrl_env* env, void* s) {
    ENV(env);
    if (E->own_stream && E->stream) { cudaStreamSynchronize(E->stream); cudaStreamDestroy(E->stream); }
    E->stream = reinterpret_cast<cudaStream_t>(s); E->own_stream = false; return 0;
}

int irrl_init(irrl_env* env) {
    ENV(env);
    if (E->initialised) return 0;
    const size_t N = (size_t)E->P.N;
    if (!E->stream) { CUDA_OK(cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking)); E->own_stream = true; }
    int rc = 0;
    rc |= dev_alloc(E, &E->S.base, N * 16); rc |= dev_alloc(E, &E->S.cmd, N * 8); rc |= dev_alloc(E, &E->S.legs, N * 64);
    rc |= dev_alloc(E, &E->S.legs2, N * 48); rc |= dev_alloc(E, &E->S.obd, N * 36); rc |= dev_alloc(E, &E->S.obd_last, N * 36);
    rc |= dev_alloc(E, &E->S.legmodel, N * 64); rc |= dev_alloc(E, &E->S.basemodel, N * 8);
    rc |= dev_alloc(E, &E->S.frame_idx, N); rc |= dev_alloc(E, &E->S.itera, N); rc |= dev_alloc(E, &E->S.ep_len, N);
    rc |= dev_alloc(E, &E->S.ep_ret, N); rc |= dev_alloc(E, &E->S.solver_sweeps, N);
    if (E->P.flag_crucial) rc |= dev_alloc(E, &E->S.meteor, N * 12); else E->S.meteor = nullptr;
    rc |= dev_alloc(E, &E->d_action, N * 12);
    {   // step outputs live in one block [ob | reward | extra | done]: a caller that lays its host buffers out the same way gets them in one copy
        unsigned char* blk = nullptr; rc |= dev_alloc(E, &blk, N * (35 + 1 + 6) * 4 + N);
        if (!rc) { E->d_ob = reinterpret_cast<float*>(blk); E->d_reward = E->d_ob + N * 35; E->d_extra = E->d_reward + N; E->d_done = reinterpret_cast<uint8_t*>(E->d_extra + N * 6); }
    } rc |= dev_alloc(E, &E->d_ep_ret, N); rc |= dev_alloc(E, &E->d_ep_len, N);
    if (rc) return rc;
    if (int r2 = ensure_pin(E, N * (35 + 12 + 1 + 6 + 2) * sizeof(float) + N)) return r2;
    // optional reference table (VEC:158-169); a missing file is only a warning there (VEC:71-73)
    if (!E->P.flag_manual_traj && !E->P.flag_manual && !E->ref_path.empty() && !E->d_ref) {
        std::vector<float> tab; int rows, cols;
        if (load_csv(E->ref_path, tab, rows, cols) == 0 && cols == 30) { if (int r3 = irrl_set_ref_traj(env, tab.data(), rows)) return r3; }
    }
    if (!E->P.flag_manual_traj && !E->P.flag_manual && !E->d_ref)
        return fail(-3, "ManualTraj: False needs a reference table: RefTraj file with 30 columns (ENV:17-21) or irrl_set_ref_traj()");
    if (E->P.flag_terrain && !E->d_terrain) {   // ENV:252-265: 500 x 20 m, 5000 x 500 samples
        if (int r4 = irrl_generate_terrain(env, nullptr, 5000, 500, 500.0, 20.0)) return r4;
    }
    launch_env_init(E->P, E->S, E->stream); CUDA_OK(cudaGetLastError());
    E->initialised = true;
    // VEC:172-182: every env is reset once during init
    StepArgs a = make_args(E, nullptr, nullptr, nullptr, nullptr, nullptr);
    launch_env_reset(a, E->stream); CUDA_OK(cudaGetLastError());
    if (E->P.flag_crucial) { launch_env_meteor(a, 1, E->stream); CUDA_OK(cudaGetLastError()); }     // ENV:608-611
    E->tick++;
    CUDA_OK(cudaStreamSynchronize(E->stream));
    return 0;
}

int irrl_get_ob_dim(irrl_env*) { return OB_DIM; }
int irrl_get_action_dim(irrl_env*) { return ACT_DIM; }
int irrl_get_extra_info_dim(irrl_env*) { return EXTRA_DIM; }
int irrl_get_origin_state_dim(irrl_env*) { return IRRL_ORIGIN_STATE_DIM; }
int irrl_get_num_envs(irrl_env* env) { irrl_env_impl* E = reinterpret_cast<irrl_env_impl*>(env); return E ? E->P.N : 0; }
const char* irrl_get_extra_info_name(irrl_env* env, int i) {
    irrl_env_impl* E = reinterpret_cast<irrl_env_impl*>(env); if (!E || i < 0 || i >= (int)E->extra_names.size()) return nullptr; return E->extra_names[i].c_str();
}

int irrl_reset(irrl_env* env, float* ob) {
    NvtxRange nvtx_("irrl_reset");
    ENV(env); NEED_INIT();
    bool dev = is_device_ptr(ob);
    StepArgs a = make_args(E, nullptr, E->P.flag_obs_filter ? nullptr : (dev ? ob : E->d_ob), nullptr, nullptr, nullptr);
    launch_env_reset(a, E->stream); CUDA_OK(cudaGetLastError());
    if (E->P.flag_crucial) { launch_env_meteor(a, 1, E->stream); CUDA_OK(cudaGetLastError()); }     // ENV:608-611
    E->tick++;
    if (E->P.flag_obs_filter) { launch_env_observe(E->P, E->S, dev ? ob : E->d_ob, E->stream); CUDA_OK(cudaGetLastError()); }
    if (!dev) return deliver(E, ob, E->d_ob, (size_t)E->P.N * OB_DIM * sizeof(float));
    return 0;
}

int irrl_observe(irrl_env* env, float* ob) {
    ENV(env); NEED_INIT();
    return probe(E, ob, OB_DIM, [&](float* d) { launch_env_observe(E->P, E->S, d, E->stream); });
}

static int step_impl(irrl_env_impl* E, const float* action, float* ob, float* reward, uint8_t* done, float* extra) {
    const size_t N = (size_t)E->P.N;
    const bool dev = is_device_ptr(action);
    if (dev) {
        if ((ob && !is_device_ptr(ob)) || (reward && !is_device_ptr(reward)) || (done && !is_device_ptr(done)) || (extra && !is_device_ptr(extra)))
            return fail(-1, "irrl_step: action is device memory, so every output must be device memory too");
        if (!reward || !done) return fail(-1, "irrl_step: reward and done are required");
        StepArgs a = make_args(E, action, E->P.flag_obs_filter ? nullptr : ob, reward, done, extra);
        if (E->P.flag_crucial) launch_env_meteor(a, 0, E->stream);
        if (E->profiling) {   // same event triplets as irrl_rollout (act slot empty): kernel time without host-side call overhead
            while (E->prof_events.size() < E->prof_cursor + 3) { cudaEvent_t ev; CUDA_OK(cudaEventCreate(&ev)); E->prof_events.push_back(ev); }
            CUDA_OK(cudaEventRecord(E->prof_events[E->prof_cursor + 0], E->stream)); CUDA_OK(cudaEventRecord(E->prof_events[E->prof_cursor + 1], E->stream));
        }
        launch_env_step(a, E->stream); CUDA_OK(cudaGetLastError());
        if (E->profiling) { CUDA_OK(cudaEventRecord(E->prof_events[E->prof_cursor + 2], E->stream)); E->prof_cursor += 3; }
        E->tick++;
        if (E->P.flag_obs_filter && ob) { launch_env_observe(E->P, E->S, ob, E->stream); CUDA_OK(cudaGetLastError()); }
        return 0;
    }
    if (!action || !reward || !done) return fail(-1, "irrl_step: action, reward and done are required");
    // host path: one H2D, one kernel, D2H of the results, one synchronisation.  Page-locked caller buffers
    // (cudaHostRegister / irrl_host_register) are DMA targets themselves; pageable ones go through pinned staging.
    unsigned char* pin = E->h_pin;
    float* p_act = reinterpret_cast<float*>(pin);
    float* p_ob = p_act + N * 12; float* p_rew = p_ob + N * 35; float* p_ext = p_rew + N; uint8_t* p_done = reinterpret_cast<uint8_t*>(p_ext + N * 6);
    // buffers registered with irrl_host_register are used in place by the kernel (zero copy); other page-locked buffers are DMA
    // targets; pageable ones go through pinned staging
    const bool outs_ok = zero_copy_mode() == 1;
    const float* z_act = mapped_alias(action, N * 48);
    float* z_ob = outs_ok && !E->P.flag_obs_filter ? mapped_alias(ob, N * 140) : nullptr; float* z_rew = outs_ok ? mapped_alias(reward, N * 4) : nullptr;
    uint8_t* z_done = outs_ok ? mapped_alias(done, N) : nullptr; float* z_ext = outs_ok ? mapped_alias(extra, N * 24) : nullptr;
    const bool pa = !z_act && is_pinned_host_ptr(action, N * 48), po = ob && !z_ob && is_pinned_host_ptr(ob, N * 140), pr = !z_rew && is_pinned_host_ptr(reward, N * 4),
               pd = !z_done && is_pinned_host_ptr(done, N), pe = extra && !z_ext && is_pinned_host_ptr(extra, N * 24);
    if (!z_act) {
        if (!pa) memcpy(p_act, action, N * 12 * sizeof(float));
        CUDA_OK(cudaMemcpyAsync(E->d_action, pa ? action : p_act, N * 12 * sizeof(float), cudaMemcpyHostToDevice, E->stream));
    }
    StepArgs a = make_args(E, z_act ? z_act : E->d_action, E->P.flag_obs_filter ? nullptr : (z_ob ? z_ob : E->d_ob), z_rew ? z_rew : E->d_reward, z_done ? z_done : E->d_done,
                           extra ? (z_ext ? z_ext : E->d_extra) : E->d_extra);
    if (E->P.flag_crucial) launch_env_meteor(a, 0, E->stream);
    launch_env_step(a, E->stream); CUDA_OK(cudaGetLastError());
    E->tick++;
    if (E->P.flag_obs_filter) { launch_env_observe(E->P, E->S, E->d_ob, E->stream); CUDA_OK(cudaGetLastError()); }
    // host buffers laid out like the device block (ob | reward | extra | done back to back, all page-locked): one copy instead of four
    const bool packed = ob && extra && po && pr && pe && pd && reward == ob + N * 35 && extra == reward + N && done == reinterpret_cast<uint8_t*>(extra + N * 6);
    if (packed) {
        CUDA_OK(cudaMemcpyAsync(ob, E->d_ob, N * (35 + 1 + 6) * sizeof(float) + N, cudaMemcpyDeviceToHost, E->stream));
    } else {
        if (ob && !z_ob) CUDA_OK(cudaMemcpyAsync(po ? ob : p_ob, E->d_ob, N * 35 * sizeof(float), cudaMemcpyDeviceToHost, E->stream));
        if (!z_rew) CUDA_OK(cudaMemcpyAsync(pr ? reward : p_rew, E->d_reward, N * sizeof(float), cudaMemcpyDeviceToHost, E->stream));
        if (extra && !z_ext) CUDA_OK(cudaMemcpyAsync(pe ? extra : p_ext, E->d_extra, N * 6 * sizeof(float), cudaMemcpyDeviceToHost, E->stream));
        if (!z_done) CUDA_OK(cudaMemcpyAsync(pd ? done : p_done, E->d_done, N, cudaMemcpyDeviceToHost, E->stream));
    }
    CUDA_OK(wait_stream(E->stream));
    if (ob && !z_ob && !po) memcpy(ob, p_ob, N * 35 * sizeof(float));
    if (!z_rew && !pr) memcpy(reward, p_rew, N * sizeof(float));
    if (extra && !z_ext && !pe) memcpy(extra, p_ext, N * 6 * sizeof(float));
    if (!z_done && !pd) memcpy(done, p_done, N);
    return 0;
}

int irrl_step(irrl_env* env, const float* action, float* ob, float* reward, uint8_t* done, float* extra) {
    NvtxRange nvtx_("irrl_step");
    ENV(env); NEED_INIT();
    return step_impl(E, action, ob, reward, done, extra);
}

int irrl_test_step(irrl_env* env, const float* action, float* ob, float* reward, uint8_t* done, float* extra) {
    // VEC:280-290: only environment 0 advances; rows 1.. of the outputs are left untouched
    ENV(env); NEED_INIT();
    if (is_device_ptr(action)) return fail(-1, "irrl_test_step takes host buffers");
    if (!action || !reward || !done) return fail(-1, "irrl_test_step: action, reward and done are required");
    EnvParams saved = E->P; E->P.N = 1;
    float ob1[35], rew1, ext1[6]; uint8_t done1;
    // stage only row 0
    CUDA_OK(cudaMemcpyAsync(E->d_action, action, 12 * sizeof(float), cudaMemcpyHostToDevice, E->stream));
    StepArgs a = make_args(E, E->d_action, E->P.flag_obs_filter ? nullptr : E->d_ob, E->d_reward, E->d_done, E->d_extra);
    if (E->P.flag_crucial) launch_env_meteor(a, 0, E->stream);
    launch_env_step(a, E->stream); cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess && E->P.flag_obs_filter) { launch_env_observe(E->P, E->S, E->d_ob, E->stream); ce = cudaGetLastError(); }
    E->P = saved; E->tick++;
    CUDA_OK(ce);
    CUDA_OK(cudaMemcpyAsync(ob1, E->d_ob, sizeof(ob1), cudaMemcpyDeviceToHost, E->stream));
    CUDA_OK(cudaMemcpyAsync(&rew1, E->d_reward, sizeof(float), cudaMemcpyDeviceToHost, E->stream));
    CUDA_OK(cudaMemcpyAsync(ext1, E->d_extra, sizeof(ext1), cudaMemcpyDeviceToHost, E->stream));
    CUDA_OK(cudaMemcpyAsync(&done1, E->d_done, 1, cudaMemcpyDeviceToHost, E->stream));
    CUDA_OK(cudaStreamSynchronize(E->stream));
    if (ob) memcpy(ob, ob1, sizeof(ob1));
    reward[0] = rew1; done[0] = done1; if (extra) memcpy(extra, ext1, sizeof(ext1));
    return 0;
}

int irrl_last_episode_stats(irrl_env* env, float* ep_return, int32_t* ep_length) {
    ENV(env); NEED_INIT();
    if (int rc = deliver(E, ep_return, E->d_ep_ret, (size_t)E->P.N * sizeof(float))) return rc;
    return deliver(E, ep_length, E->d_ep_len, (size_t)E->P.N * sizeof(int));
}
int irrl_running_episode_stats(irrl_env* env, float* ep_return, int32_t* ep_length, int clear) {
    ENV(env); NEED_INIT();
    if (int rc = deliver(E, ep_return, E->S.ep_ret, (size_t)E->P.N * sizeof(float))) return rc;
    if (int rc = deliver(E, ep_length, E->S.ep_len, (size_t)E->P.N * sizeof(int))) return rc;
    if (clear) { CUDA_OK(cudaMemsetAsync(E->S.ep_ret, 0, (size_t)E->P.N * sizeof(float), E->stream)); CUDA_OK(cudaMemsetAsync(E->S.ep_len, 0, (size_t)E->P.N * sizeof(int), E->stream)); }
    return 0;
}

int irrl_set_seed(irrl_env* env, int seed) { ENV(env); E->P.seed = (uint32_t)seed; return 0; }
int irrl_close(irrl_env* env) { ENV(env); if (E->stream) CUDA_OK(cudaStreamSynchronize(E->stream)); return 0; }
int irrl_set_simulation_time_step(irrl_env* env, double dt) { ENV(env); E->P.sim_dt = (float)dt; E->sim_dt_d = dt; E->P.loop_count = int(E->control_dt_d / dt + 1e-10); return 0; }
int irrl_set_control_time_step(irrl_env* env, double dt) {
    ENV(env); E->P.control_dt = (float)dt; E->control_dt_d = dt; E->P.control_dt_d = dt; E->P.loop_count = int(dt / E->sim_dt_d + 1e-10);
    E->P.disturb_every = int(E->period_d / dt * 10.0); E->P.meteor_every = int(5.0 * E->period_d / dt);     // ENV:746, 731 evaluate control_dt_ at every step
    if (E->d_ref) {   // frame_len = int(max_time / control_dt_) (ENV:539) follows the control step; the table must still be long enough
        const int frame_len = int(E->max_time_d / dt);
        if (E->P.ref_rows / 2 <= frame_len + 10) return fail(-3, "setControlTimeStep: the reference table is too short for this control step (ENV:538-539, 571)");
        E->P.frame_len = frame_len;
    }
    return 0;
}
int irrl_curriculum_update(irrl_env* env) { ENV(env); return 0; }
int irrl_start_recording_video(irrl_env* env, const char*) { ENV(env); return 0; }
int irrl_stop_recording_video(irrl_env* env) { ENV(env); return 0; }
int irrl_show_window(irrl_env* env) { ENV(env); return 0; }
int irrl_hide_window(irrl_env* env) { ENV(env); return 0; }

int irrl_get_state(irrl_env* env, float* out) { ENV(env); NEED_INIT(); return probe(E, out, STATE_DIM, [&](float* d) { launch_env_get_state(E->P, E->S, d, E->stream); }); }
int irrl_set_state(irrl_env* env, const float* in) {
    ENV(env); NEED_INIT();
    size_t n = (size_t)E->P.N * STATE_DIM;
    const float* src = in;
    if (!is_device_ptr(in)) { if (int rc = ensure_scratch(E, n)) return rc; CUDA_OK(cudaMemcpyAsync(E->d_scratch, in, n * sizeof(float), cudaMemcpyHostToDevice, E->stream)); src = E->d_scratch; }
    launch_env_set_state(E->P, E->S, src, E->stream); CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(E->stream));
    return 0;
}
int irrl_set_tick(irrl_env* env, uint32_t tick) { ENV(env); E->tick = tick; return 0; }
uint32_t irrl_get_tick(irrl_env* env) { irrl_env_impl* E = reinterpret_cast<irrl_env_impl*>(env); return E ? E->tick : 0; }

int irrl_is_terminal_state(irrl_env* env, uint8_t* terminal) {
    ENV(env); NEED_INIT();
    std::vector<float> s((size_t)E->P.N * STATE_DIM);
    if (int rc = irrl_get_state(env, s.data())) return rc;
    std::vector<uint8_t> t(E->P.N);
    for (int i = 0; i < E->P.N; ++i) {
        const float* x = &s[(size_t)i * STATE_DIM]; float z = x[2];
        if (!E->terrain_h.empty()) {   // height above the terrain under the trunk (same triangulation as the kernels)
            const EnvParams& P = E->P; int nx = P.terrain_nx, ny = P.terrain_ny;
            float fx = (x[0] - P.terrain_cx) / P.terrain_dx + 0.5f * (nx - 1), fy = (x[1] - P.terrain_cy) / P.terrain_dy + 0.5f * (ny - 1);
            fx = std::min(std::max(fx, 0.f), (float)(nx - 1)); fy = std::min(std::max(fy, 0.f), (float)(ny - 1));
            int ii = std::min((int)fx, nx - 2), jj = std::min((int)fy, ny - 2); float uu = fx - ii, vv = fy - jj;
            const float* hh = &E->terrain_h[(size_t)ii * ny + jj]; float h00 = hh[0], h01 = hh[1], h10 = hh[ny], h11 = hh[ny + 1];
            z -= (uu >= vv) ? h00 + (h10 - h00) * uu + (h11 - h10) * vv : h00 + (h11 - h01) * uu + (h01 - h00) * vv;
        }
        t[i] = (z < 0.15f || z > 0.65f || x[109 + 31] < 0.5f) ? 1 : 0;   // ENV:1560
    }
    if (is_device_ptr(terminal)) { CUDA_OK(cudaMemcpy(terminal, t.data(), t.size(), cudaMemcpyHostToDevice)); } else memcpy(terminal, t.data(), t.size());
    return 0;
}

static int state_slice(irrl_env* env, float* out, int width, const std::vector<std::pair<int, int>>& spans) {
    irrl_env_impl* E = reinterpret_cast<irrl_env_impl*>(env);
    std::vector<float> s((size_t)E->P.N * STATE_DIM), o((size_t)E->P.N * width);
    if (int rc = irrl_get_state(env, s.data())) return rc;
    for (int i = 0; i < E->P.N; ++i) { int c = 0; for (auto& sp : spans) for (int k = 0; k < sp.second; ++k) o[(size_t)i * width + c++] = s[(size_t)i * STATE_DIM + sp.first + k]; }
    if (is_device_ptr(out)) { CUDA_OK(cudaMemcpy(out, o.data(), o.size() * sizeof(float), cudaMemcpyHostToDevice)); } else memcpy(out, o.data(), o.size() * sizeof(float));
    return 0;
}
int irrl_origin_state(irrl_env* env, float* out) { ENV(env); NEED_INIT(); return state_slice(env, out, 41, {{0, 19}, {19, 18}, {105, 4}}); }          // ENV:1323
int irrl_reference_state(irrl_env* env, float* out) { ENV(env); NEED_INIT(); return state_slice(env, out, 24, {{67, 12}, {79, 12}}); }                  // ENV:1343
int irrl_get_joint_effort(irrl_env* env, float* out) { ENV(env); NEED_INIT(); return state_slice(env, out, 12, {{179, 12}}); }                          // ENV:1354
int irrl_get_generalized_force(irrl_env* env, float* out) {
    ENV(env); NEED_INIT();
    std::vector<float> je((size_t)E->P.N * 12), o((size_t)E->P.N * 18, 0.f);
    if (int rc = state_slice(env, je.data(), 12, {{179, 12}})) return rc;
    for (int i = 0; i < E->P.N; ++i) for (int k = 0; k < 12; ++k) o[(size_t)i * 18 + 6 + k] = je[(size_t)i * 12 + k];                                   // ggff.head(6) = 0
    if (is_device_ptr(out)) { CUDA_OK(cudaMemcpy(out, o.data(), o.size() * sizeof(float), cudaMemcpyHostToDevice)); } else memcpy(out, o.data(), o.size() * sizeof(float));
    return 0;
}
int irrl_get_inverse_mass_matrix(irrl_env* env, float* out) { ENV(env); NEED_INIT(); return probe(E, out, 324, [&](float* d) { launch_env_probe(E->P, E->S, nullptr, d, nullptr, E->stream); }); }
int irrl_get_nonlinear(irrl_env* env, float* out) { ENV(env); NEED_INIT(); return probe(E, out, 18, [&](float* d) { launch_env_probe(E->P, E->S, nullptr, nullptr, d, E->stream); }); }
int irrl_get_mass_matrix(irrl_env* env, float* out) { ENV(env); NEED_INIT(); return probe(E, out, 324, [&](float* d) { launch_env_probe(E->P, E->S, d, nullptr, nullptr, E->stream); }); }

int irrl_set_contact_coefficient(irrl_env* env, const float* coeff) {
    ENV(env); NEED_INIT();
    std::vector<float> c((size_t)E->P.N * 3), bm((size_t)E->P.N * 8);
    if (is_device_ptr(coeff)) { CUDA_OK(cudaMemcpy(c.data(), coeff, c.size() * sizeof(float), cudaMemcpyDeviceToHost)); } else memcpy(c.data(), coeff, c.size() * sizeof(float));
    CUDA_OK(cudaStreamSynchronize(E->stream));
    CUDA_OK(cudaMemcpy(bm.data(), E->S.basemodel, bm.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int i = 0; i < E->P.N; ++i) { bm[(size_t)i * 8 + 4] = c[(size_t)i * 3]; bm[(size_t)i * 8 + 5] = c[(size_t)i * 3 + 1]; bm[(size_t)i * 8 + 6] = c[(size_t)i * 3 + 2]; }
    CUDA_OK(cudaMemcpy(E->S.basemodel, bm.data(), bm.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}
// ENV:1423-1436: position and radius of the (first) attack sphere
int irrl_get_sphere_info(irrl_env* env, float* out) {
    ENV(env); NEED_INIT();
    if (!E->P.flag_crucial) return fail(-3, "Please make sure the [Flag_Crucial] is True");   // ENV:1434
    const size_t N = (size_t)E->P.N;
    std::vector<float> m(N * 12), o(N * 4);
    CUDA_OK(cudaStreamSynchronize(E->stream)); CUDA_OK(cudaMemcpy(m.data(), E->S.meteor, m.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < N; ++i) { o[4 * i] = m[12 * i]; o[4 * i + 1] = m[12 * i + 1]; o[4 * i + 2] = m[12 * i + 2]; o[4 * i + 3] = m[12 * i + 7]; }
    if (is_device_ptr(out)) { CUDA_OK(cudaMemcpy(out, o.data(), o.size() * 4, cudaMemcpyHostToDevice)); } else memcpy(out, o.data(), o.size() * 4);
    return 0;
}
// test access to the meteor state: [N,9] = p(3) v(3) mode radius mass (host memory)
int irrl_get_meteor(irrl_env* env, float* out) {
    ENV(env); NEED_INIT();
    if (!E->P.flag_crucial) return fail(-3, "irrl_get_meteor: Crutial is False");
    const size_t N = (size_t)E->P.N; std::vector<float> m(N * 12);
    CUDA_OK(cudaStreamSynchronize(E->stream)); CUDA_OK(cudaMemcpy(m.data(), E->S.meteor, m.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < N; ++i) for (int k = 0; k < 9; ++k) out[9 * i + k] = m[12 * i + k];
    return 0;
}
int irrl_set_meteor(irrl_env* env, const float* in) {
    ENV(env); NEED_INIT();
    if (!E->P.flag_crucial) return fail(-3, "irrl_set_meteor: Crutial is False");
    const size_t N = (size_t)E->P.N; std::vector<float> m(N * 12, 0.f);
    for (size_t i = 0; i < N; ++i) for (int k = 0; k < 9; ++k) m[12 * i + k] = in[9 * i + k];
    CUDA_OK(cudaStreamSynchronize(E->stream)); CUDA_OK(cudaMemcpy(E->S.meteor, m.data(), m.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}
int irrl_get_model_params(irrl_env* env, float* out) {
    ENV(env); NEED_INIT();
    const int N = E->P.N; std::vector<float> bm((size_t)N * 8), lm((size_t)N * 64), o((size_t)N * 94);
    CUDA_OK(cudaStreamSynchronize(E->stream));
    CUDA_OK(cudaMemcpy(bm.data(), E->S.basemodel, bm.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(lm.data(), E->S.legmodel, lm.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; ++i) {
        float* r = &o[(size_t)i * 94]; const float* b = &bm[(size_t)i * 8];
        r[0] = b[4]; r[1] = b[5]; r[2] = b[6];
        r[3] = b[0]; r[4] = b[1]; r[5] = b[2]; r[6] = b[3]; r[7] = r[8] = r[9] = 0.f;
        for (int l = 0; l < 4; ++l) {
            const float* m = &lm[((size_t)i * 4 + l) * 16]; float sx = (l < 2) ? 1.f : -1.f, sy = (l & 1) ? 1.f : -1.f;
            const float offs[3][3] = {{E->P.off1x * sx, E->P.off1y * sy, 0.f}, {0.f, E->P.off2y * sy, 0.f}, {0.f, 0.f, m[3]}};
            for (int b3 = 0; b3 < 3; ++b3) { float* q = r + 3 + 7 * (1 + 3 * l + b3); q[0] = m[b3]; for (int a = 0; a < 3; ++a) { q[1 + a] = m[4 + 4 * b3 + a]; q[4 + a] = offs[b3][a]; } }
        }
    }
    if (is_device_ptr(out)) { CUDA_OK(cudaMemcpy(out, o.data(), o.size() * sizeof(float), cudaMemcpyHostToDevice)); } else memcpy(out, o.data(), o.size() * sizeof(float));
    return 0;
}

int irrl_integrate(irrl_env* env, const float* tau, float* contact_out) {
    ENV(env); NEED_INIT();
    const size_t N = (size_t)E->P.N;
    const float* dtau = tau;
    if (!is_device_ptr(tau)) { CUDA_OK(cudaMemcpyAsync(E->d_action, tau, N * 12 * sizeof(float), cudaMemcpyHostToDevice, E->stream)); dtau = E->d_action; }
    bool cdev = is_device_ptr(contact_out);
    if (int rc = ensure_scratch(E, N * 16)) return rc;
    launch_env_integrate(E->P, E->S, dtau, contact_out ? (cdev ? contact_out : E->d_scratch) : nullptr, E->stream); CUDA_OK(cudaGetLastError());
    if (contact_out && !cdev) return deliver(E, contact_out, E->d_scratch, N * 16 * sizeof(float));
    CUDA_OK(cudaStreamSynchronize(E->stream));
    return 0;
}
int irrl_get_solver_sweeps(irrl_env* env, int32_t* out) { ENV(env); NEED_INIT(); return deliver(E, out, E->S.solver_sweeps, (size_t)E->P.N * sizeof(int)); }

// ---- heightfield terrain.  The reference builds a RaiSim HeightMap from TerrainProperties (frequency 1, zScale 0.1, 500 x 20 m,
// 5000 x 500 samples, 3 fractal octaves, lacunarity 2, gain 0.25; ENV:252-265).  RaiSim's noise generator is not available, so
// "perlin" here is a fractal value-noise with the same parameters (new specification, unpinned); "stairs" is the stair course of
// BASELINE.json config 5 (rise / run along +x).  Any other table can be supplied with irrl_set_heightfield.
static inline float lattice(int ix, int iy, uint32_t seed) {
    uint32_t h = (uint32_t)ix * 0x8da6b343u ^ (uint32_t)iy * 0xd8163841u ^ seed * 0xcb1ab31fu;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
    return (float)(h & 0xffffff) / 8388608.0f - 1.0f;
}
static inline float value_noise(float x, float y, uint32_t seed) {
    int ix = (int)std::floor(x), iy = (int)std::floor(y); float fx = x - ix, fy = y - iy;
    float sx = fx * fx * (3 - 2 * fx), sy = fy * fy * (3 - 2 * fy);
    float a = lattice(ix, iy, seed), b = lattice(ix + 1, iy, seed), c2 = lattice(ix, iy + 1, seed), d = lattice(ix + 1, iy + 1, seed);
    return (a + (b - a) * sx) + ((c2 + (d - c2) * sx) - (a + (b - a) * sx)) * sy;
}
int irrl_set_heightfield(irrl_env* env, const float* heights, int nx, int ny, double x_size, double y_size, double cx, double cy) {
    ENV(env);
    if (!heights || nx < 2 || ny < 2) return fail(-1, "heightfield needs at least 2 x 2 samples");
    E->terrain_h.resize((size_t)nx * ny);
    if (is_device_ptr(heights)) { CUDA_OK(cudaMemcpy(E->terrain_h.data(), heights, E->terrain_h.size() * 4, cudaMemcpyDeviceToHost)); } else memcpy(E->terrain_h.data(), heights, E->terrain_h.size() * 4);
    float* d = nullptr; CUDA_OK(cudaMalloc((void**)&d, E->terrain_h.size() * 4)); E->allocs.push_back(d);
    CUDA_OK(cudaMemcpy(d, E->terrain_h.data(), E->terrain_h.size() * 4, cudaMemcpyHostToDevice));
    E->d_terrain = d; E->P.terrain = d; E->P.terrain_nx = nx; E->P.terrain_ny = ny; E->P.terrain_cx = (float)cx; E->P.terrain_cy = (float)cy;
    E->P.terrain_dx = (float)(x_size / (nx - 1)); E->P.terrain_dy = (float)(y_size / (ny - 1)); E->P.flag_terrain = 1;
    return 0;
}
int irrl_generate_terrain(irrl_env* env, const char* kind, int nx, int ny, double x_size, double y_size) {
    ENV(env);
    std::string k = kind ? kind : E->terrain_kind;
    std::vector<float> h((size_t)nx * ny);
    const double dx = x_size / (nx - 1), dy = y_size / (ny - 1);
    for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) {
        double x = -x_size / 2 + i * dx, yy = -y_size / 2 + j * dy; float v = 0.f;
        if (k == "stairs") { if (x >= E->stair_start) v = (float)(E->stair_rise * (std::floor((x - E->stair_start) / E->stair_run) + 1.0)); }
        else if (k == "flat") v = 0.f;
        else {   // "perlin": TerrainProperties of ENV:254-263
            float amp = 1.f, freq = 1.f, sum = 0.f;
            for (int o = 0; o < 3; ++o) { sum += amp * value_noise((float)x * freq, (float)yy * freq, (uint32_t)(E->terrain_seed + 31 * o)); amp *= 0.25f; freq *= 2.0f; }
            v = 0.1f * sum;
        }
        h[(size_t)i * ny + j] = v;
    }
    return irrl_set_heightfield(env, h.data(), nx, ny, x_size, y_size, 0.0, 0.0);
}
int irrl_get_heightfield(irrl_env* env, float* heights, int* nx, int* ny, double* x_size, double* y_size, double* cx, double* cy) {
    ENV(env);
    if (E->terrain_h.empty()) return fail(-3, "no heightfield is set (Terrain: False)");
    if (nx) *nx = E->P.terrain_nx; if (ny) *ny = E->P.terrain_ny;
    if (x_size) *x_size = (double)E->P.terrain_dx * (E->P.terrain_nx - 1); if (y_size) *y_size = (double)E->P.terrain_dy * (E->P.terrain_ny - 1);
    if (cx) *cx = E->P.terrain_cx; if (cy) *cy = E->P.terrain_cy;
    if (heights) memcpy(heights, E->terrain_h.data(), E->terrain_h.size() * 4);
    return 0;
}

int irrl_set_ref_traj(irrl_env* env, const float* table, int rows) {
    ENV(env);
    if (!table || rows <= 0) return fail(-1, "empty reference table");
    {   // reset() draws its start row from [0, rows/2 - frame_len - 10) (ENV:571): a shorter table would index before row 0
        const int frame_len = int(E->max_time_d / E->control_dt_d);
        if (rows / 2 <= frame_len + 10) return fail(-3, "reference table too short: rows/2 = " + std::to_string(rows / 2) + " must exceed max_time/control_dt + 10 = " + std::to_string(frame_len + 10) + " (ENV:538-539, 571)");
    }
    float* d = nullptr; CUDA_OK(cudaMalloc((void**)&d, (size_t)rows * 30 * sizeof(float))); E->allocs.push_back(d);
    CUDA_OK(cudaMemcpy(d, table, (size_t)rows * 30 * sizeof(float), is_device_ptr(table) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    E->d_ref = d; E->P.ref = d; E->P.ref_rows = rows; E->P.frame_max = rows / 2; E->P.frame_len = int(E->max_time_d / E->control_dt_d);   // ENV:538-539 (double, like the reference)
    return 0;
}

// ------------------------------------------------------------------ measured FP32 (non-tensor) peak: register-resident FMA chains
__global__ void fp32_peak_kernel(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 123.456f) out[0] = a0;
}
int irrl_measure_fp32_peak(int device, double* tflops) {
    CUDA_OK(cudaSetDevice(device));
    float* d; CUDA_OK(cudaMalloc((void**)&d, 4));
    cudaDeviceProp prop; CUDA_OK(cudaGetDeviceProperties(&prop, device));
    const int iters = 4096, threads = 256, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0); fp32_peak_kernel<<<blocks, threads>>>(d, iters); cudaEventRecord(e1); CUDA_OK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 64.0 * iters * (double)threads * blocks;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (tflops) *tflops = best;
    return 0;
}

// ------------------------------------------------------------------ policy
static void bind_weights(irrl_policy_impl* Pn) {
    const float* p = Pn->d_params; PolicyWeights& W = Pn->W;
    const int in[4] = {35, 48, 35, 48};
    for (int i = 0; i < 4; ++i) { W.wx[i] = p; p += in[i] * 192; W.wh[i] = p; p += 48 * 192; W.b[i] = p; p += 192; }
    W.vf_w = p; p += 48; W.vf_b = p; p += 1; W.pi_w = p; p += 48 * 12; W.pi_b = p; p += 12; W.logstd = p; p += 12; /* q head (unused by act) follows */
    for (int i = 0; i < 4; ++i) { W.wcat[i] = Pn->d_derived + (size_t)i * (96 * 192 + 192); W.bperm[i] = W.wcat[i] + 96 * 192; }
    W.tcblob = Pn->d_tcblob;
}
// gate-interleaved, zero-padded [x ; h] weight blocks: new column (u/2)*8 + (u%2)*4 + g  <-  original column g*48 + u
static int upload_derived(irrl_policy_impl* Pn, const float* host_params) {
    std::vector<float> d((size_t)4 * (96 * 192 + 192), 0.f);
    const float* p = host_params; const int in[4] = {35, 48, 35, 48};
    for (int i = 0; i < 4; ++i) {
        const float* wx = p; p += in[i] * 192; const float* wh = p; p += 48 * 192; const float* b = p; p += 192;
        float* o = d.data() + (size_t)i * (96 * 192 + 192);
        for (int g = 0; g < 4; ++g) for (int uu = 0; uu < 48; ++uu) {
            int src = g * 48 + uu, dst = (uu / 2) * 8 + (uu % 2) * 4 + g;
            for (int k = 0; k < in[i]; ++k) o[(size_t)k * 192 + dst] = wx[(size_t)k * 192 + src];
            for (int k = 0; k < 48; ++k) o[(size_t)(in[i] + k) * 192 + dst] = wh[(size_t)k * 192 + src];
            o[96 * 192 + dst] = b[src];
        }
    }
    CUDA_OK(cudaMemcpy(Pn->d_derived, d.data(), d.size() * sizeof(float), cudaMemcpyHostToDevice));
    {   // tensor-core layout (policy_tc_kernels.cu): tower 0 = pi cells + mean head, tower 1 = V cells + value head
        std::vector<unsigned char> blob((size_t)2 * TC_BLOB_BYTES);
        const float* q = host_params; const float* wx[4]; const float* wh[4]; const float* bb[4];
        for (int i = 0; i < 4; ++i) { wx[i] = q; q += in[i] * 192; wh[i] = q; q += 48 * 192; bb[i] = q; q += 192; }
        const float* vf_w = q; q += 48; const float* vf_b = q; q += 1; const float* pi_w = q; q += 48 * 12; const float* pi_b = q; q += 12; const float* logstd = q;
        pack_tc_blob(wx[0], wh[0], bb[0], wx[1], wh[1], bb[1], pi_w, pi_b, 12, logstd, blob.data());
        pack_tc_blob(wx[2], wh[2], bb[2], wx[3], wh[3], bb[3], vf_w, vf_b, 1, nullptr, blob.data() + TC_BLOB_BYTES);
        CUDA_OK(cudaMemcpy(Pn->d_tcblob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    }
    return 0;
}
int irrl_policy_create(int device, const float* params, irrl_policy** out) {
    if (!out || !params) return fail(-1, "null argument");
    int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(-2, "no CUDA device: this library has no CPU fallback"); }
    CUDA_OK(cudaSetDevice(device));
    irrl_policy_impl* Pn = new irrl_policy_impl(); Pn->device = device;
    CUDA_OK(cudaMalloc((void**)&Pn->d_params, IRRL_POLICY_NUM_PARAMS * sizeof(float)));
    CUDA_OK(cudaMalloc((void**)&Pn->d_derived, (size_t)4 * (96 * 192 + 192) * sizeof(float)));
    CUDA_OK(cudaMalloc((void**)&Pn->d_tcblob, (size_t)2 * TC_BLOB_BYTES));
    bind_weights(Pn);
    *out = reinterpret_cast<irrl_policy*>(Pn);
    return irrl_policy_set_params(*out, params);
}
int irrl_policy_set_params(irrl_policy* pol, const float* params) {
    irrl_policy_impl* Pn = reinterpret_cast<irrl_policy_impl*>(pol); if (!Pn) return fail(-1, "null policy");
    CUDA_OK(cudaSetDevice(Pn->device));
    CUDA_OK(cudaMemcpy(Pn->d_params, params, IRRL_POLICY_NUM_PARAMS * sizeof(float), is_device_ptr(params) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    std::vector<float> hp(IRRL_POLICY_NUM_PARAMS);
    CUDA_OK(cudaMemcpy(hp.data(), Pn->d_params, hp.size() * sizeof(float), cudaMemcpyDeviceToHost));
    return upload_derived(Pn, hp.data());
}
// CPU-callable: the compact robot model read from a URDF file as 40 floats (order of model_to_array); path NULL -> the built-in constants
int irrl_parse_urdf(const char* path, float* out40) {
    if (!out40) return fail(-1, "null argument");
    EnvParams P{}; model_defaults(P); P.joint_damping = 0.01f;
    if (path) { CompactModel M; std::string uerr; if (int rc = read_urdf(path, M, uerr)) return fail(rc, std::string(path) + ": " + uerr); apply_urdf(P, M); }
    model_to_array(P, out40); return 0;
}
int irrl_policy_set_act_path(int mode) {
    if (mode < 0 || mode > 2) return fail(-1, "irrl_policy_set_act_path: mode must be 0 (auto), 1 (fp32 FMA kernel) or 2 (tcgen05 kernel)");
    g_act_path = mode; return 0;
}
int irrl_tc_timeline(int enable, long long* out16) { tc_timeline(enable, out16); return 0; }
int irrl_tc_mma_rate(int n, int reps, unsigned layout_type, unsigned lbo, unsigned sbo, unsigned kadv, long long* cycles2) {
    long long* d = nullptr; CUDA_OK(cudaMalloc((void**)&d, 16));
    int rc = launch_tc_mma_rate(n, reps, layout_type, lbo, sbo, kadv, d, 0);
    if (rc) { cudaFree(d); return fail(-1, "irrl_tc_mma_rate: launch failed"); }
    CUDA_OK(cudaGetLastError()); CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(cycles2, d, 16, cudaMemcpyDeviceToHost)); cudaFree(d);
    return 0;
}
int irrl_tc_gemm_probe(const float* a, const float* b, float* d, int k, int n, int variant) {
    if (!a || !b || !d) return fail(-1, "null argument");
    float *da = nullptr, *db = nullptr, *dd = nullptr;
    CUDA_OK(cudaMalloc((void**)&da, 128 * k * 4)); CUDA_OK(cudaMalloc((void**)&db, (size_t)n * k * 4)); CUDA_OK(cudaMalloc((void**)&dd, (size_t)128 * n * 4));
    CUDA_OK(cudaMemcpy(da, a, 128 * k * 4, cudaMemcpyHostToDevice)); CUDA_OK(cudaMemcpy(db, b, (size_t)n * k * 4, cudaMemcpyHostToDevice));
    int rc = launch_tc_gemm_probe(da, db, dd, k, n, variant, 0);
    if (rc) { cudaFree(da); cudaFree(db); cudaFree(dd); return fail(-1, "irrl_tc_gemm_probe: k must be a multiple of 8 in [8,64], n a multiple of 16 in [16,256]"); }
    CUDA_OK(cudaGetLastError()); CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(d, dd, (size_t)128 * n * 4, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return 0;
}
void irrl_policy_destroy(irrl_policy* pol) {
    irrl_policy_impl* Pn = reinterpret_cast<irrl_policy_impl*>(pol); if (!Pn) return;
    cudaSetDevice(Pn->device); cudaDeviceSynchronize();
    cudaFree(Pn->d_params); cudaFree(Pn->d_derived); cudaFree(Pn->d_tcblob); cudaFree(Pn->d_obs); cudaFree(Pn->d_state); cudaFree(Pn->d_action); cudaFree(Pn->d_done);
    if (Pn->h_pin) cudaFreeHost(Pn->h_pin);
    delete Pn;
}
int irrl_policy_act(irrl_policy* pol, void* cuda_stream, int n, const float* obs, const uint8_t* done, float* state, float* action, float* clipped,
                    float* value, float* neglogp, int deterministic, uint32_t seed, uint32_t env_offset, uint32_t tick) {
    NvtxRange nvtx_("irrl_policy_act");
    irrl_policy_impl* Pn = reinterpret_cast<irrl_policy_impl*>(pol); if (!Pn) return fail(-1, "null policy");
    if (!obs || !state || !action || !value || !neglogp || n <= 0) return fail(-1, "irrl_policy_act: null argument");
    CUDA_OK(cudaSetDevice(Pn->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    ActArgs a; a.W = Pn->W; a.N = n; a.deterministic = deterministic; a.seed = seed; a.env_offset = env_offset; a.tick = tick; a.mean = nullptr; a.obs_store = nullptr; a.done_store = nullptr;
    const size_t N = (size_t)n;
    // every pointer is classified on its own: e.g. the LSTM state may stay resident on the device while
    // observations / actions travel through host memory (the reference's Runner keeps `states` opaque, ppo2.py:520)
    const bool h_obs = !is_device_ptr(obs), h_done = done && !is_device_ptr(done), h_state = !is_device_ptr(state), h_act = !is_device_ptr(action),
               h_clip = clipped && !is_device_ptr(clipped), h_val = !is_device_ptr(value), h_nlp = !is_device_ptr(neglogp);
    const bool any_host = h_obs || h_done || h_state || h_act || h_clip || h_val || h_nlp;
    if (any_host && n > Pn->cap) {
        cudaFree(Pn->d_obs); cudaFree(Pn->d_state); cudaFree(Pn->d_action); cudaFree(Pn->d_done);
        if (Pn->h_pin) cudaFreeHost(Pn->h_pin);
        CUDA_OK(cudaMalloc((void**)&Pn->d_obs, N * 35 * 4)); CUDA_OK(cudaMalloc((void**)&Pn->d_state, N * 384 * 4));
        CUDA_OK(cudaMalloc((void**)&Pn->d_action, N * (12 + 12 + 1 + 1) * 4));      // one block [action | clipped | value | neglogp]: one copy for callers laid out the same way
        Pn->d_clipped = Pn->d_action + N * 12; Pn->d_value = Pn->d_clipped + N * 12; Pn->d_nlp = Pn->d_value + N; CUDA_OK(cudaMalloc((void**)&Pn->d_done, N));
        CUDA_OK(cudaMallocHost((void**)&Pn->h_pin, N * (35 + 384 + 12 + 12 + 1 + 1) * 4 + N));
        Pn->cap = n;
    }
    float* p_obs = reinterpret_cast<float*>(Pn->h_pin); float* p_state = p_obs + N * 35; float* p_act = p_state + N * 384; float* p_clip = p_act + N * 12;
    float* p_val = p_clip + N * 12; float* p_nlp = p_val + N; uint8_t* p_done = reinterpret_cast<uint8_t*>(p_nlp + N);
    // zero copy for buffers registered with irrl_host_register (see step_impl); the LSTM state is read-modify-write and
    // always staged when it lives on the host
    const bool outs_ok = zero_copy_mode() == 1;
    const float* z_obs = h_obs ? mapped_alias(obs, N * 140) : nullptr; const uint8_t* z_done = h_done ? mapped_alias(done, N) : nullptr;
    float* z_act = h_act && outs_ok ? mapped_alias(action, N * 48) : nullptr; float* z_clip = h_clip && outs_ok ? mapped_alias(clipped, N * 48) : nullptr;
    float* z_val = h_val && outs_ok ? mapped_alias(value, N * 4) : nullptr; float* z_nlp = h_nlp && outs_ok ? mapped_alias(neglogp, N * 4) : nullptr;
    const bool q_obs = h_obs && !z_obs && is_pinned_host_ptr(obs, N * 140), q_done = h_done && !z_done && is_pinned_host_ptr(done, N), q_state = h_state && is_pinned_host_ptr(state, N * 1536),
               q_act = h_act && !z_act && is_pinned_host_ptr(action, N * 48), q_clip = h_clip && !z_clip && is_pinned_host_ptr(clipped, N * 48),
               q_val = h_val && !z_val && is_pinned_host_ptr(value, N * 4), q_nlp = h_nlp && !z_nlp && is_pinned_host_ptr(neglogp, N * 4);
    if (h_obs && !z_obs) { if (!q_obs) memcpy(p_obs, obs, N * 35 * 4); CUDA_OK(cudaMemcpyAsync(Pn->d_obs, q_obs ? obs : p_obs, N * 35 * 4, cudaMemcpyHostToDevice, st)); }
    if (h_state) { if (!q_state) memcpy(p_state, state, N * 384 * 4); CUDA_OK(cudaMemcpyAsync(Pn->d_state, q_state ? state : p_state, N * 384 * 4, cudaMemcpyHostToDevice, st)); }
    if (h_done && !z_done) { if (!q_done) memcpy(p_done, done, N); CUDA_OK(cudaMemcpyAsync(Pn->d_done, q_done ? done : p_done, N, cudaMemcpyHostToDevice, st)); }
    a.obs = h_obs ? (z_obs ? z_obs : Pn->d_obs) : obs; a.done = done ? (h_done ? (z_done ? z_done : Pn->d_done) : done) : nullptr; a.state = h_state ? Pn->d_state : state;
    a.action = h_act ? (z_act ? z_act : Pn->d_action) : action; a.clipped = clipped ? (h_clip ? (z_clip ? z_clip : Pn->d_clipped) : clipped) : nullptr;
    a.value = h_val ? (z_val ? z_val : Pn->d_value) : value; a.neglogp = h_nlp ? (z_nlp ? z_nlp : Pn->d_nlp) : neglogp;
    launch_lstm_act(a, st); CUDA_OK(cudaGetLastError());
    if (!any_host) return 0;
    if (h_state) CUDA_OK(cudaMemcpyAsync(q_state ? state : p_state, Pn->d_state, N * 384 * 4, cudaMemcpyDeviceToHost, st));
    const bool packed = q_act && q_clip && q_val && q_nlp && clipped == action + N * 12 && value == clipped + N * 12 && neglogp == value + N;
    if (packed) {
        CUDA_OK(cudaMemcpyAsync(action, Pn->d_action, N * (12 + 12 + 1 + 1) * 4, cudaMemcpyDeviceToHost, st));
    } else {
        if (h_act && !z_act) CUDA_OK(cudaMemcpyAsync(q_act ? action : p_act, Pn->d_action, N * 12 * 4, cudaMemcpyDeviceToHost, st));
        if (h_clip && !z_clip) CUDA_OK(cudaMemcpyAsync(q_clip ? clipped : p_clip, Pn->d_clipped, N * 12 * 4, cudaMemcpyDeviceToHost, st));
        if (h_val && !z_val) CUDA_OK(cudaMemcpyAsync(q_val ? value : p_val, Pn->d_value, N * 4, cudaMemcpyDeviceToHost, st));
        if (h_nlp && !z_nlp) CUDA_OK(cudaMemcpyAsync(q_nlp ? neglogp : p_nlp, Pn->d_nlp, N * 4, cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(wait_stream(st));
    if (h_state && !q_state) memcpy(state, p_state, N * 384 * 4);
    if (h_act && !z_act && !q_act) memcpy(action, p_act, N * 12 * 4);
    if (h_clip && !z_clip && !q_clip) memcpy(clipped, p_clip, N * 12 * 4);
    if (h_val && !z_val && !q_val) memcpy(value, p_val, N * 4);
    if (h_nlp && !z_nlp && !q_nlp) memcpy(neglogp, p_nlp, N * 4);
    return 0;
}

int irrl_rollout(irrl_env* env, irrl_policy* pol, int T, const irrl_rollout_buffers* b, int deterministic) {
    NvtxRange nvtx_("irrl_rollout");
    ENV(env); NEED_INIT();
    irrl_policy_impl* Pn = reinterpret_cast<irrl_policy_impl*>(pol); if (!Pn || !b) return fail(-1, "null argument");
    if (E->P.flag_obs_filter) return fail(-3, "irrl_rollout does not support ObsFilter: True");
    const size_t N = (size_t)E->P.N;
    for (const void* p : {(const void*)b->obs, (const void*)b->actions, (const void*)b->values, (const void*)b->neglogps, (const void*)b->rewards,
                          (const void*)b->dones, (const void*)b->cur_obs, (const void*)b->cur_done, (const void*)b->state})
        if (!is_device_ptr(p)) return fail(-1, "irrl_rollout: all buffers must be device memory");
    // profiling never blocks: events are appended per step and only read back by irrl_get_profile
    if (E->profiling) while (E->prof_events.size() < E->prof_cursor + 3 * (size_t)T) { cudaEvent_t ev; CUDA_OK(cudaEventCreate(&ev)); E->prof_events.push_back(ev); }
    for (int t = 0; t < T; ++t) {
        // mb_obs / mb_dones hold the inputs of model.step (ppo2.py:520-526)
        ActArgs a; a.W = Pn->W; a.N = (int)N; a.deterministic = deterministic; a.seed = E->P.seed; a.env_offset = E->P.env_offset; a.tick = E->tick; a.mean = nullptr;
        a.obs_store = b->obs + (size_t)t * N * 35; a.done_store = b->dones + (size_t)t * N;     // stored by the act kernel itself
        a.obs = b->cur_obs; a.done = b->cur_done; a.state = b->state; a.action = b->actions + (size_t)t * N * 12; a.clipped = E->d_action;
        a.value = b->values + (size_t)t * N; a.neglogp = b->neglogps + (size_t)t * N;
        if (E->profiling) CUDA_OK(cudaEventRecord(E->prof_events[E->prof_cursor + 0], E->stream));
        launch_lstm_act(a, E->stream);
        if (E->profiling) CUDA_OK(cudaEventRecord(E->prof_events[E->prof_cursor + 1], E->stream));
        StepArgs s = make_args(E, E->d_action, b->cur_obs, b->rewards + (size_t)t * N, b->cur_done, nullptr);
        if (b->ep_return && b->ep_length) { s.ep_ret_out = b->ep_return + (size_t)t * N; s.ep_len_out = b->ep_length + (size_t)t * N; }
        if (E->P.flag_crucial) launch_env_meteor(s, 0, E->stream);
        launch_env_step(s, E->stream);
        if (E->profiling) { CUDA_OK(cudaEventRecord(E->prof_events[E->prof_cursor + 2], E->stream)); E->prof_cursor += 3; }
        E->tick++;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- one host entry per control step: model.step (run_bp_v5.py:178-185) -> np.clip (ppo2.py:529-531) -> env.step (RaisimGymVecEnv.py:26-52)
// The environments are cut into chunks that travel through separate streams: the host->device copy of chunk k+1 and the
// device->host copy of chunk k-1 overlap the two kernels of chunk k, so the PCIe time of a step hides behind its compute.
static int batch_copy(void** dst, void** src, size_t* bytes, int n, cudaMemcpyKind kind, cudaStream_t st) {
    static bool use_batch = [] { const char* e = getenv("IRRL_BATCH_COPY"); return !e || atoi(e) != 0; }();
    if (use_batch && n > 1) {
        cudaMemcpyAttributes attr{}; attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream; attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
        size_t idx = 0, fail = 0;
        cudaError_t e = cudaMemcpyBatchAsync(dst, src, bytes, (size_t)n, &attr, &idx, 1, &fail, st);
        if (e == cudaSuccess) return 0;
        cudaGetLastError(); use_batch = false;                      // not supported by this driver / stream: plain copies from now on
    }
    for (int i = 0; i < n; ++i) CUDA_OK(cudaMemcpyAsync(dst[i], src[i], bytes[i], kind, st));
    return 0;
}
int irrl_act_step(irrl_env* env, irrl_policy* pol, const irrl_act_step_io* io, int deterministic, uint32_t act_tick, int chunks) {
    NvtxRange nvtx_("irrl_act_step");
    ENV(env); NEED_INIT();
    irrl_policy_impl* Pn = reinterpret_cast<irrl_policy_impl*>(pol); if (!Pn || !io) return fail(-1, "null argument");
    if (E->P.flag_obs_filter) return fail(-3, "irrl_act_step does not support ObsFilter: True (use irrl_policy_act + irrl_step)");
    if (!io->obs || !io->state || !io->action || !io->value || !io->neglogp || !io->next_obs || !io->reward || !io->next_done) return fail(-1, "irrl_act_step: null buffer");
    if (!is_device_ptr(io->state)) return fail(-1, "irrl_act_step: the LSTM state must be device memory (it stays resident between steps)");
    const size_t N = (size_t)E->P.N;
    const bool host = !is_device_ptr(io->obs);
    const void* all[] = {io->obs, io->done, io->action, io->clipped, io->value, io->neglogp, io->next_obs, io->reward, io->next_done, io->extra};
    const size_t all_bytes[] = {N * 140, N, N * 48, N * 48, N * 4, N * 4, N * 140, N * 4, N, N * 24};
    for (int i = 0; i < 10; ++i) {
        if (!all[i]) continue;
        if (host ? !is_pinned_host_ptr(all[i], all_bytes[i]) : !is_device_ptr(all[i]))
            return fail(-1, host ? "irrl_act_step: host buffers must be page-locked (irrl_host_register / cudaHostAlloc): the copies run asynchronously on several streams"
                                 : "irrl_act_step: obs is device memory, so every buffer must be device memory");
    }
    // device images of the host buffers: [action | clipped | value | neglogp] next to the env's own [ob | reward | extra | done] block
    if (host && !E->d_fused_act) { void* q = nullptr; CUDA_OK(cudaMalloc(&q, N * 26 * sizeof(float))); E->allocs.push_back(q); E->d_fused_act = (float*)q; }
    // measured (scripts/r2_e2e.py): the kernels are latency-bound below ~8k environments (a half-size launch takes as long as a full
    // one) and every extra stream hop costs ~5 us, so the batch is cut only from 12k environments on (2 chunks)
    int K = chunks > 0 ? chunks : (N >= 24576 ? 4 : (N >= 12288 ? 2 : 1));        // measured (scripts/r2_e2e.py): 16384 envs 262 / 280 us with 2 / 4 chunks, 32768 envs 461 / 397 us
    if (!host || E->P.flag_crucial) K = 1;
    K = std::min(K, 8);
    size_t per = ((N + K - 1) / K + 127) / 128 * 128;      // chunk boundaries on multiples of 128 environments (the tile of the tensor-core act kernel)
    while (E->n_chunk_streams < std::max(K, 1)) {
        CUDA_OK(cudaStreamCreateWithFlags(&E->chunk_stream[E->n_chunk_streams], cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&E->chunk_done[E->n_chunk_streams], cudaEventDisableTiming));
        CUDA_OK(cudaStreamCreateWithFlags(&E->side_stream[E->n_chunk_streams], cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&E->act_done[E->n_chunk_streams], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&E->side_done[E->n_chunk_streams], cudaEventDisableTiming)); E->n_chunk_streams++;
    }
    if (!E->fork_ev) CUDA_OK(cudaEventCreateWithFlags(&E->fork_ev, cudaEventDisableTiming));
    float *d_act = host ? E->d_fused_act : io->action, *d_clip = host ? E->d_fused_act + N * 12 : (io->clipped ? io->clipped : E->d_action),
          *d_val = host ? E->d_fused_act + N * 24 : io->value, *d_nlp = host ? E->d_fused_act + N * 25 : io->neglogp;
    float *d_ob = host ? E->d_ob : io->next_obs, *d_rew = host ? E->d_reward : io->reward, *d_ext = host ? E->d_extra : io->extra; uint8_t* d_done = host ? E->d_done : io->next_done;
    const float* d_obs_in = host ? E->d_ob : io->obs; const uint8_t* d_done_in = host ? (io->done ? E->d_done : nullptr) : io->done;
    // host buffers laid out like the device images (fused_host_step in vec_env.py does that): one copy per block instead of one per array
    const bool in_place = io->obs == io->next_obs && (!io->done || io->done == io->next_done);
    const bool env_block = host && io->extra && io->reward == io->next_obs + N * 35 && io->extra == io->reward + N && io->next_done == reinterpret_cast<uint8_t*>(io->extra + N * 6);
    const bool act_block = host && io->clipped && io->clipped == io->action + N * 12 && io->value == io->clipped + N * 12 && io->neglogp == io->value + N;
    if (E->P.flag_crucial) { StepArgs m = make_args(E, nullptr, nullptr, nullptr, nullptr, nullptr); launch_env_meteor(m, 0, E->stream); }
    if (K > 1) CUDA_OK(cudaEventRecord(E->fork_ev, E->stream));
    // the act results (action, clipped, value, neglogp: 104 of the 273 bytes per env that go home) are copied on a side stream as soon as the act
    // kernel of their chunk has finished, i.e. behind its step kernel, instead of after it (measured: 4096 envs 137.6 -> 128.4 us per step,
    // 8192: 184 -> 167, 16384: 270 -> 257; IRRL_ACT_SIDE=0 switches that off)
    static const int side_mode = [] { const char* e = getenv("IRRL_ACT_SIDE"); return e ? atoi(e) : 2; }();      // 0 off, 1 only with chunks, 2 always (default)
    const bool side = side_mode != 0 && host && (K > 1 || side_mode == 2);
    int used = 0;
    for (size_t lo = 0; lo < N; lo += per, ++used) {
        const size_t hi = std::min(N, lo + per), n = hi - lo;
        cudaStream_t st = K == 1 ? E->stream : E->chunk_stream[used];
        if (K > 1) CUDA_OK(cudaStreamWaitEvent(st, E->fork_ev, 0));
        if (host) {
            if (K == 1 && env_block && in_place && io->done) {     // [ob | reward | extra | done] in one copy (reward / extra ride along: 28 of 169 bytes per env)
                CUDA_OK(cudaMemcpyAsync(E->d_ob, io->obs, N * 168 + N, cudaMemcpyHostToDevice, st));
            } else {
                void* dst[2] = {(void*)(E->d_ob + lo * 35), (void*)(E->d_done + lo)}; void* src[2] = {(void*)(io->obs + lo * 35), (void*)(io->done ? io->done + lo : nullptr)};
                size_t by[2] = {n * 140, n};
                if (int rc = batch_copy(dst, src, by, io->done ? 2 : 1, cudaMemcpyHostToDevice, st)) return rc;
            }
        }
        ActArgs a; a.W = Pn->W; a.N = (int)n; a.deterministic = deterministic; a.seed = E->P.seed; a.env_offset = E->P.env_offset + (uint32_t)lo; a.tick = act_tick; a.n_total = (int)N;
        a.mean = nullptr; a.obs_store = nullptr; a.done_store = nullptr;
        a.obs = d_obs_in + lo * 35; a.done = d_done_in ? d_done_in + lo : nullptr; a.state = io->state + lo * 384;
        a.action = d_act + lo * 12; a.clipped = d_clip + lo * 12; a.value = d_val + lo; a.neglogp = d_nlp + lo;
        launch_lstm_act(a, st);
        if (side) {
            CUDA_OK(cudaEventRecord(E->act_done[used], st)); CUDA_OK(cudaStreamWaitEvent(E->side_stream[used], E->act_done[used], 0));
            void* dst[4]; void* src[4]; size_t by[4]; int c = 0;
            auto add = [&](void* h, void* d, size_t b) { if (h) { dst[c] = h; src[c] = d; by[c] = b; ++c; } };
            add(io->action + lo * 12, d_act + lo * 12, n * 48); if (io->clipped) add(io->clipped + lo * 12, d_clip + lo * 12, n * 48);
            add(io->value + lo, d_val + lo, n * 4); add(io->neglogp + lo, d_nlp + lo, n * 4);
            if (K == 1 && env_block && act_block) { CUDA_OK(cudaMemcpyAsync(io->action, E->d_fused_act, N * 26 * sizeof(float), cudaMemcpyDeviceToHost, E->side_stream[used])); }
            else if (int rc = batch_copy(dst, src, by, c, cudaMemcpyDeviceToHost, E->side_stream[used])) return rc;
            CUDA_OK(cudaEventRecord(E->side_done[used], E->side_stream[used]));
        }
        StepArgs s = make_args(E, d_clip, d_ob, d_rew, d_done, d_ext ? d_ext : E->d_extra);
        s.r_begin = (int)lo; s.r_end = (int)hi;
        launch_env_step(s, st);
        if (host) {
            if (K == 1 && env_block && act_block) {
                if (!side) CUDA_OK(cudaMemcpyAsync(io->action, E->d_fused_act, N * 26 * sizeof(float), cudaMemcpyDeviceToHost, st));
                CUDA_OK(cudaMemcpyAsync(io->next_obs, E->d_ob, N * 168 + N, cudaMemcpyDeviceToHost, st));
            } else {
                void* dst[8]; void* src[8]; size_t by[8]; int c = 0;
                auto add = [&](void* h, void* d, size_t b) { if (h) { dst[c] = h; src[c] = d; by[c] = b; ++c; } };
                if (!side) {
                    add(io->action + lo * 12, d_act + lo * 12, n * 48); if (io->clipped) add(io->clipped + lo * 12, d_clip + lo * 12, n * 48);
                    add(io->value + lo, d_val + lo, n * 4); add(io->neglogp + lo, d_nlp + lo, n * 4);
                }
                add(io->next_obs + lo * 35, d_ob + lo * 35, n * 140); add(io->reward + lo, d_rew + lo, n * 4);
                if (io->extra) add(io->extra + lo * 6, E->d_extra + lo * 6, n * 24); add(io->next_done + lo, d_done + lo, n);
                if (int rc = batch_copy(dst, src, by, c, cudaMemcpyDeviceToHost, st)) return rc;
            }
        }
        if (K > 1) { CUDA_OK(cudaEventRecord(E->chunk_done[used], st)); CUDA_OK(cudaStreamWaitEvent(E->stream, E->chunk_done[used], 0)); }
        if (side) CUDA_OK(cudaStreamWaitEvent(E->stream, E->side_done[used], 0));
    }
    CUDA_OK(cudaGetLastError());
    E->tick++;
    if (host) CUDA_OK(cudaStreamSynchronize(E->stream));
    return 0;
}

int irrl_set_profiling(irrl_env* env, int on) {
    ENV(env);
    if (E->stream) CUDA_OK(cudaStreamSynchronize(E->stream));
    E->profiling = on != 0; E->prof_cursor = 0; E->prof_act_ms = E->prof_step_ms = 0; E->prof_count = 0; return 0;
}
int irrl_get_profile(irrl_env* env, double* act_ms_total, double* step_ms_total, int64_t* launches_each) {
    ENV(env);
    if (E->stream) CUDA_OK(cudaStreamSynchronize(E->stream));
    E->prof_act_ms = E->prof_step_ms = 0; E->prof_count = 0;
    for (size_t i = 0; i + 2 < E->prof_cursor; i += 3) {
        float m1 = 0, m2 = 0;
        CUDA_OK(cudaEventElapsedTime(&m1, E->prof_events[i], E->prof_events[i + 1]));
        CUDA_OK(cudaEventElapsedTime(&m2, E->prof_events[i + 1], E->prof_events[i + 2]));
        E->prof_act_ms += m1; E->prof_step_ms += m2; E->prof_count++;
    }
    if (act_ms_total) *act_ms_total = E->prof_act_ms;
    if (step_ms_total) *step_ms_total = E->prof_step_ms;
    if (launches_each) *launches_each = E->prof_count;
    return 0;
}

int irrl_lstm_seq_fwd(void* cuda_stream, int T, int K, int n_env, const float* xw, const float* wh, const float* c0, const float* h0, const float* keep,
                      float* gates, float* Cs, float* Hs, const float* bias, float* HM) {
    launch_lstm_seq_fwd(T, K, n_env, xw, wh, c0, h0, keep, gates, Cs, Hs, bias, HM, reinterpret_cast<cudaStream_t>(cuda_stream)); CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_lstm_seq_bwd(void* cuda_stream, int T, int K, int n_env, const float* dH, const float* wh, const float* c0, const float* keep, const float* gates,
                      const float* Cs, float* dz, float* db_part) {
    launch_lstm_seq_bwd(T, K, n_env, dH, wh, c0, keep, gates, Cs, dz, db_part, reinterpret_cast<cudaStream_t>(cuda_stream)); CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_lstm_seq_ctas(int n_env) { return lstm_seq_ctas(n_env); }
int irrl_lstm_seq_set_path(int path) { return lstm_seq_set_path(path); }
int irrl_proj_rows(void* cuda_stream, int T, int K, int n_env, const float* X, int x_cols, int x_has_tower, const float* W, int w_trans, float* Y, int n_out) {
    NvtxRange nvtx_("irrl_proj_rows");
    if (!X || !W || !Y || T <= 0 || K <= 0 || n_env <= 0) return fail(-1, "irrl_proj_rows: bad argument");
    int rc = -1;
    if (w_trans && n_out == 48 && x_cols == 192 && x_has_tower) rc = launch_projT_rows_tc(X, W, Y, T, K, n_env, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (!w_trans && n_out == 192) rc = launch_proj_rows_tc(X, x_cols, x_has_tower, W, Y, T, K, n_env, reinterpret_cast<cudaStream_t>(cuda_stream));      // tcgen05 where the shape fits a TMEM tile
    if (rc == -1) rc = launch_proj_rows(X, x_cols, x_has_tower, W, w_trans, Y, n_out, T, K, n_env, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (rc == -1) return fail(-1, "irrl_proj_rows: unsupported shape (x_cols <= 40 or 48 with n_out 192; x_cols 192 with n_out 48)");
    if (rc) return fail(rc, "irrl_proj_rows: kernel configuration failed");
    CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_proj_rows_set_path(int path) { return proj_rows_set_path(path); }
int irrl_scale_unless_one(void* cuda_stream, float* x, long long n, const float* scale) {
    if (!x || !scale || n < 0 || (n & 3) || (reinterpret_cast<uintptr_t>(x) & 15)) return fail(-1, "irrl_scale_unless_one: bad argument (n % 4 == 0, 16-byte aligned)");
    if (n) launch_scale_unless_one(x, n, scale, reinterpret_cast<cudaStream_t>(cuda_stream));
    CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_ppo_head_loss_ctas(int T, int n_env) { return ppo_head_loss_ctas((long long)T * n_env); }
int irrl_ppo_head_loss(void* cuda_stream, int T, int n_env, const float* H1, const float* pi_w, const float* pi_b, const float* vf_w, const float* vf_b, const float* logstd,
                       const float* actions, const float* adv, const float* returns, const float* old_values, const float* old_neglogp,
                       float cliprange, float vf_coef, float inv_count, float* dH, float* G, float* partial) {
    NvtxRange nvtx_("irrl_ppo_head_loss");
    if (!H1 || !pi_w || !pi_b || !vf_w || !vf_b || !logstd || !actions || !adv || !returns || !old_values || !old_neglogp || !dH || !G || !partial || T <= 0 || n_env <= 0)
        return fail(-1, "irrl_ppo_head_loss: bad argument");
    launch_ppo_head_loss(H1, pi_w, pi_b, vf_w, vf_b, logstd, actions, adv, returns, old_values, old_neglogp, dH, G, partial, cliprange, vf_coef, inv_count, T, n_env,
                         reinterpret_cast<cudaStream_t>(cuda_stream));
    CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_gram2_rows_ctas(int T, int K, int n_env) { return gram2_rows_tc_ctas(T, n_env, K); }
int irrl_gram2_rows(void* cuda_stream, int T, int K, int n_env, const float* X, int x_cols, int x_has_tower, const float* Hs, const float* h0, const float* keep, const float* D,
                    float* partial) {
    NvtxRange nvtx_("irrl_gram2_rows");
    if (!X || !Hs || !h0 || !keep || !D || !partial || T <= 0 || K <= 0 || n_env <= 0) return fail(-1, "irrl_gram2_rows: bad argument");
    const int rc = launch_gram2_rows_tc(X, x_cols, x_has_tower, Hs, h0, keep, D, partial, T, K, n_env, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (rc == -1) return fail(-1, "irrl_gram2_rows: x_cols must be in 1..48");
    if (rc == -3) return fail(-3, "irrl_gram2_rows: needs n_env % 4 == 0 and 16-byte aligned tensors (use irrl_gram_rows)");
    if (rc) return fail(rc, "irrl_gram2_rows: kernel configuration failed");
    CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_gram_rows_ctas(int T, int K, int n_env) { return gram_rows_ctas(T, n_env, K); }
int irrl_gram_rows(void* cuda_stream, int T, int K, int n_env, const float* X, int x_cols, int x_has_tower, const float* D, float* partial) {
    NvtxRange nvtx_("irrl_gram_rows");
    if (!X || !D || !partial || T <= 0 || K <= 0 || n_env <= 0) return fail(-1, "irrl_gram_rows: bad argument");
    const int rc = launch_gram_rows(X, x_cols, x_has_tower, D, partial, T, K, n_env, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (rc == -1) return fail(-1, "irrl_gram_rows: x_cols must be in 1..48");
    if (rc) return fail(rc, "irrl_gram_rows: kernel configuration failed");
    CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_lstm_pw_fwd(void* cuda_stream, int rows, int n_env, const float* z, const float* c_prev_masked, const float* keep_next, float* gates,
                     float* c_out, float* h_out, float* hm_next, float* cm_next) {
    launch_lstm_pw_fwd(rows, n_env, z, c_prev_masked, keep_next, gates, c_out, h_out, hm_next, cm_next, reinterpret_cast<cudaStream_t>(cuda_stream));
    CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_lstm_pw_bwd(void* cuda_stream, int rows, int n_env, const float* dh_out, const float* carry_h, const float* carry_c, const float* keep_up,
                     const float* gates, const float* c, const float* c_prev_masked, float* dz, float* dcm_prev) {
    launch_lstm_pw_bwd(rows, n_env, dh_out, carry_h, carry_c, keep_up, gates, c, c_prev_masked, dz, dcm_prev, reinterpret_cast<cudaStream_t>(cuda_stream));
    CUDA_OK(cudaGetLastError()); return 0;
}
int irrl_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(-1, "irrl_host_register: null argument");
    if (is_pinned_host_ptr(ptr, bytes)) return 0;
    CUDA_OK(cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    void* dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, ptr, 0) == cudaSuccess && dev) {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        g_pinned[reinterpret_cast<uintptr_t>(ptr)] = PinnedRange{bytes, static_cast<char*>(dev)};
    } else cudaGetLastError();
    return 0;
}
int irrl_host_unregister(void* ptr) {
    if (!ptr) return 0;
    { std::lock_guard<std::mutex> lk(g_pinned_mu); g_pinned.erase(reinterpret_cast<uintptr_t>(ptr)); }
    cudaError_t e = cudaHostUnregister(ptr); if (e != cudaSuccess) cudaGetLastError(); return 0;
}

int irrl_gae(void* cuda_stream, int T, int n, const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
             const uint8_t* last_dones, float gamma, float lam, float* adv, float* returns) {
    launch_gae(rewards, values, dones, last_values, last_dones, adv, returns, T, n, gamma, lam, reinterpret_cast<cudaStream_t>(cuda_stream));
    CUDA_OK(cudaGetLastError()); return 0;
}

}  // extern "C"
