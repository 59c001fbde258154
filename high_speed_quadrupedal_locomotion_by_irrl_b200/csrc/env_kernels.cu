// sm_100a kernels of the bp5 environment: step (PD + 8 physics substeps + obs + reward + done + auto-reset),
// reset, observe, probes (M, M^-1, h) and flat state get/set.  One robot = one quad of lanes (env_device.cuh).
//
// Replaces, for the hot path, VectorizedEnvironment::step (VEC:268-278), perAgentStep (VEC:352-372),
// ENVIRONMENT::step/reset/observe (ENV:692-809, 547-635, 1248-1262) and RaiSim's world.integrate().
#include <cstdlib>
#include "env_device.cuh"
#include "env_kernels.h"

namespace irrl {

constexpr int BLOCK = 64;   // 16 robots per CTA: small CTAs spread 4096 robots (512 warps) over all 148 SMs

// ------------------------------------------------------------------ register-resident env state of one quad
struct EnvRegs {
    Base b;
    f3 q, qd, torque_last, jref, jdref, ptl, eeref, tau_applied;
    float contact_flag, impulse_norm;
    float cmd[3], cmdf[3];
    float t0; int frame_idx, itera, ep_len; float ep_ret;
    LegModel lm; BaseModel bm;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }

// What the substep loop needs: base state, joint state, torque filter memory, previous joint targets, the per-env model.
__device__ __forceinline__ void load_env_hot(const DevState& S, int r, int leg, EnvRegs& e) {
    const float* bp = S.base + (size_t)r * 16;
    float4 b0 = ld4(bp), b1 = ld4(bp + 4), b2 = ld4(bp + 8), b3 = ld4(bp + 12);
    e.b.p = mk(b0.x, b0.y, b0.z); e.b.qw = b0.w; e.b.qx = b1.x; e.b.qy = b1.y; e.b.qz = b1.z;
    e.b.v = mk(b1.w, b2.x, b2.y); e.b.w = mk(b2.z, b2.w, b3.x);
    const float* lp = S.legs + ((size_t)r * 4 + leg) * 16;
    float4 l0 = ld4(lp), l1 = ld4(lp + 4), l2 = ld4(lp + 8);
    e.q = mk(l0.x, l0.y, l0.z); e.qd = mk(l0.w, l1.x, l1.y); e.torque_last = mk(l1.z, l1.w, l2.x);
    const float* l2p = S.legs2 + ((size_t)r * 4 + leg) * 12;
    float4 m0 = ld4(l2p);
    e.ptl = mk(m0.x, m0.y, m0.z);
    const float* mp = S.legmodel + ((size_t)r * 4 + leg) * 16;
    float4 g0 = ld4(mp), g1 = ld4(mp + 4), g2 = ld4(mp + 8), g3 = ld4(mp + 12);
    e.lm.m1 = g0.x; e.lm.m2 = g0.y; e.lm.m3 = g0.z; e.lm.knee_z = g0.w;
    e.lm.com1 = mk(g1.x, g1.y, g1.z); e.lm.com2 = mk(g2.x, g2.y, g2.z); e.lm.com3 = mk(g3.x, g3.y, g3.z);
    e.lm.sx = (leg < 2) ? 1.f : -1.f; e.lm.sy = (leg & 1) ? 1.f : -1.f;
    const float* bmp = S.basemodel + (size_t)r * 8;
    float4 h0 = ld4(bmp), h1 = ld4(bmp + 4);
    e.bm.m0 = h0.x; e.bm.com0 = mk(h0.y, h0.z, h0.w); e.bm.mu = h1.x; e.bm.rest = h1.y; e.bm.thr = h1.z;
}
// Everything else (clock, command, references, episode counters): only read after the substeps.  The step kernel loads it there, so
// that these ~35 values are not live (in registers or spilled) across the hot loop.
__device__ __forceinline__ void load_env_cold(const DevState& S, int r, int leg, EnvRegs& e) {
    e.t0 = S.base[(size_t)r * 16 + 13];
    const float* cp = S.cmd + (size_t)r * 8;
    float4 c0 = ld4(cp), c1 = ld4(cp + 4);
    e.cmd[0] = c0.x; e.cmd[1] = c0.y; e.cmd[2] = c0.z; e.cmdf[0] = c0.w; e.cmdf[1] = c1.x; e.cmdf[2] = c1.y;
    const float* lp = S.legs + ((size_t)r * 4 + leg) * 16;
    float4 l2 = ld4(lp + 8), l3 = ld4(lp + 12);
    e.jref = mk(l2.y, l2.z, l2.w); e.jdref = mk(l3.x, l3.y, l3.z);
    const float* l2p = S.legs2 + ((size_t)r * 4 + leg) * 12;
    float4 m0 = ld4(l2p), m1 = ld4(l2p + 4), m2 = ld4(l2p + 8);
    e.eeref = mk(m0.w, m1.x, m1.y); e.tau_applied = mk(m1.z, m1.w, m2.x);
    e.contact_flag = m2.y; e.impulse_norm = m2.z;
    e.frame_idx = S.frame_idx[r]; e.itera = S.itera[r]; e.ep_len = S.ep_len[r]; e.ep_ret = S.ep_ret[r];
}
__device__ __forceinline__ void load_env(const DevState& S, int r, int leg, EnvRegs& e) { load_env_hot(S, r, leg, e); load_env_cold(S, r, leg, e); }

__device__ __forceinline__ void store_env(const DevState& S, int r, int leg, const EnvRegs& e) {
    if (leg == 0) {
        float* bp = S.base + (size_t)r * 16;
        st4(bp, e.b.p.x, e.b.p.y, e.b.p.z, e.b.qw); st4(bp + 4, e.b.qx, e.b.qy, e.b.qz, e.b.v.x);
        st4(bp + 8, e.b.v.y, e.b.v.z, e.b.w.x, e.b.w.y); st4(bp + 12, e.b.w.z, e.t0, 0.f, 0.f);
        float* cp = S.cmd + (size_t)r * 8;
        st4(cp, e.cmd[0], e.cmd[1], e.cmd[2], e.cmdf[0]); st4(cp + 4, e.cmdf[1], e.cmdf[2], 0.f, 0.f);
        S.frame_idx[r] = e.frame_idx; S.itera[r] = e.itera; S.ep_len[r] = e.ep_len; S.ep_ret[r] = e.ep_ret;
    }
    float* lp = S.legs + ((size_t)r * 4 + leg) * 16;
    st4(lp, e.q.x, e.q.y, e.q.z, e.qd.x); st4(lp + 4, e.qd.y, e.qd.z, e.torque_last.x, e.torque_last.y);
    st4(lp + 8, e.torque_last.z, e.jref.x, e.jref.y, e.jref.z); st4(lp + 12, e.jdref.x, e.jdref.y, e.jdref.z, 0.f);
    float* l2p = S.legs2 + ((size_t)r * 4 + leg) * 12;
    st4(l2p, e.ptl.x, e.ptl.y, e.ptl.z, e.eeref.x); st4(l2p + 4, e.eeref.y, e.eeref.z, e.tau_applied.x, e.tau_applied.y);
    st4(l2p + 8, e.tau_applied.z, e.contact_flag, e.impulse_norm, 0.f);
}

__device__ __forceinline__ float cur_time(const EnvParams& P, const EnvRegs& e) { return fmaf((float)e.frame_idx, P.control_dt, e.t0); }
// table row of frame_idx: the reference indexes its Eigen table unchecked (ENV:1670); here a run past the last row (long episodes,
// step() without the forced reset) holds the last row instead of reading out of bounds
__device__ __forceinline__ int ref_row(const EnvParams& P, int frame_idx) { return min(max(frame_idx, 0), P.ref_rows - 1); }
// phase of leg `leg` in its gait cycle, fmod(t + phase * T, T) / T (ENV:1172-1181, 1523-1527), reduced in double: in fp32 the clock
// t ~ 1.5 s carries 1e-7 s, which the contact-reward shaping (slope 2 * 2 pi / lam) turns into 1e-5 of the reward exponent
__device__ __forceinline__ float gait_phase(const EnvParams& P, const EnvRegs& e, int leg) {
#ifdef EXP_F32_PHASE
    float rp = fmaf((float)e.frame_idx, P.control_dt, e.t0) + P.phase[leg] * P.period; return fmodf(rp, P.period) / P.period;
#else
    const double t = (double)e.t0 + (double)e.frame_idx * P.control_dt_d;
    const double x = t * P.inv_period_d + (double)P.phase[leg];       // reciprocal precomputed on the host: no fp64 division subroutine
    return (float)(x - floor(x));
#endif
}
__device__ __forceinline__ f3 nominal_q(const EnvParams& P, int leg) { return mk((leg & 1) ? P.abad : -P.abad, -0.78f, 1.57f); }   // ENV:317-322

// ------------------------------------------------------------------ observation (ENV:956-1004); writes obDouble_ to HBM
// ob29_31 / bodyvel outputs are replicated in the quad.
struct ObsOut { float ob29, ob30, ob31; f3 blin, bang; float oq[3], oqd[3], ph3, ph4, om[3]; };   // o* / ph* / om: the obDouble_ entries this lane wrote
// Quad exchanges of update_observation name only the four lanes of the robot in their mask: the function also runs inside reset_env,
// i.e. under a branch that is uniform per quad but divergent per warp.
__device__ __forceinline__ unsigned quad_mask() { return 0xFu << (threadIdx.x & 28); }
__device__ __forceinline__ float quad_get(unsigned qm, float v, int src) { return __shfl_sync(qm, v, src, 4); }
// word `w` of the Philox block that lane `src` of the quad holds (w, src differ per destination lane: one shuffle per word, then a pick)
__device__ __forceinline__ uint32_t quad_word(unsigned qm, uint4 blk, int src, int w) {
    const uint32_t a = __shfl_sync(qm, blk.x, src, 4), b = __shfl_sync(qm, blk.y, src, 4), c = __shfl_sync(qm, blk.z, src, 4), d = __shfl_sync(qm, blk.w, src, 4);
    return w == 0 ? a : (w == 1 ? b : (w == 2 ? c : d));
}
__device__ __forceinline__ void update_observation(const EnvParams& P, const DevState& S, int r, int gid, int leg, uint32_t tick,
                                                   uint32_t pbase, const EnvRegs& e, ObsOut& o) {
    float* obd = S.obd + (size_t)r * 36;
    float nq[3] = {0.f, 0.f, 0.f}, nqd[3] = {0.f, 0.f, 0.f};
    float gp[3] = {0.f, 0.f, 0.f}, go[3] = {0.f, 0.f, 0.f};
    if (P.noise_flag != 0.f) {
        // joint j = 3 leg + k lives in Philox block j/4, word j%4 (same draws as the oracle).  The quad computes every block ONCE:
        // lane l holds block min(l, 2) of the joint-angle and of the joint-rate noise and one Gaussian block (even lanes posture,
        // odd lanes angular rate); the words are handed to the lanes that need them by shuffles (was: up to 6 blocks per lane).
        const int blk = min(leg, 2); const unsigned qm = quad_mask();
        const uint4 qa = philox(P.seed, gid, tick, pbase + P_OBS_Q0 + blk), da = philox(P.seed, gid, tick, pbase + P_OBS_QD0 + blk);
        float g4[4]; gauss4(P.seed, gid, tick, pbase + ((leg & 1) ? P_OBS_OMEGA : P_OBS_POSTURE), g4);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = 3 * leg + k;
            nq[k] = usym(quad_word(qm, qa, j >> 2, j & 3)) * P.joint_noise * P.noise_flag;          // ENV:979
            nqd[k] = usym(quad_word(qm, da, j >> 2, j & 3)) * P.joint_vel_noise * P.noise_flag;     // ENV:983
            gp[k] = quad_get(qm, g4[k], 0); go[k] = quad_get(qm, g4[k], 1);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.oq[k] = nq[k] + comp(e.q, k); o.oqd[k] = nqd[k] + comp(e.qd, k); obd[5 + 3 * leg + k] = o.oq[k]; obd[17 + 3 * leg + k] = o.oqd[k]; }
    f3 bx, by, bz; quat_cols(e.b.qw, e.b.qx, e.b.qy, e.b.qz, bx, by, bz);
    o.ob29 = bx.z + (gp[0] * P.posture_sigma) * P.noise_flag;      // third row of R  ENV:994-996
    o.ob30 = by.z + (gp[1] * P.posture_sigma) * P.noise_flag;
    o.ob31 = bz.z + (gp[2] * P.posture_sigma) * P.noise_flag;
    o.blin = mk(dot(bx, e.b.v), dot(by, e.b.v), dot(bz, e.b.v));   // R^T v   ENV:999
    o.bang = mk(dot(bx, e.b.w), dot(by, e.b.w), dot(bz, e.b.w));   // R^T w   ENV:1000
    o.om[0] = o.bang.x + P.noise_flag * (go[0] * P.omega_sigma);   // ENV:1001-1003
    o.om[1] = o.bang.y + P.noise_flag * (go[1] * P.omega_sigma);
    o.om[2] = o.bang.z + P.noise_flag * (go[2] * P.omega_sigma);
    o.ph3 = 0.f; o.ph4 = 0.f;
    if (leg == 0) {
        if (P.flag_manual || P.flag_manual_traj) {                 // ENV:964-968
            float t = cur_time(P, e);
            sincosf(2.f * IRRL_PI_REF * t / P.period, &o.ph3, &o.ph4);      // one range reduction for both (same values as sinf / cosf)
        } else { const float* row = P.ref + (size_t)ref_row(P, e.frame_idx) * 30; o.ph3 = row[25]; o.ph4 = row[26]; }   // ENV:972
        obd[0] = 0.f; obd[1] = 0.f; obd[2] = 0.f;                  // obDouble_.setZero  ENV:960
        obd[3] = o.ph3; obd[4] = o.ph4;
        obd[29] = o.ob29; obd[30] = o.ob30; obd[31] = o.ob31;
        obd[32] = o.om[0]; obd[33] = o.om[1]; obd[34] = o.om[2];
    }
}

// ------------------------------------------------------------------ command + gait reference (ENV:1010-1109, 1756-1890)
__device__ __forceinline__ void command_obs_update(const EnvParams& P, const DevState& S, int r, int gid, int leg, uint32_t tick,
                                                   uint32_t purpose, bool flag_reset, EnvRegs& e) {
    if (P.flag_manual) return;                                     // ENV:1013-1019
    float* obd = S.obd + (size_t)r * 36;
    if (P.flag_manual_traj) {
        uint4 rr = philox(P.seed, gid, tick, purpose);
        float temp = u01(rr.x);
        if (temp < 0.5f / (P.max_time / P.control_dt) || flag_reset) {   // ENV:1028
            temp = u01(rr.y);
            float u3 = u01(rr.z);
            // (t < 0.2 zeroing loop is a by-value no-op, ENV:1039-1045; t <= 0.2 falls through to the last else)
            if (0.2f < temp && temp <= 0.7f) e.cmd[0] = u3 * P.Vx_max + (1.0f - u3) * P.Vx_min;            // ENV:1046-1055
            else if (0.7f < temp && temp <= 0.85f) e.cmd[1] = u3 * P.Vy_max + (1.0f - u3) * P.Vy_min;      // ENV:1058-1067
            else e.cmd[2] = u3 * P.omega_max + (1.0f - u3) * P.omega_min;                                   // ENV:1068-1077
        }
        if (flag_reset) { e.cmdf[0] = e.cmd[0]; e.cmdf[1] = e.cmd[1]; e.cmdf[2] = e.cmd[2]; }               // ENV:1080-1085
        else {
#pragma unroll
            for (int i = 0; i < 3; ++i) e.cmdf[i] = e.cmdf[i] * P.cmd_update + e.cmd[i] * (1.f - P.cmd_update);   // ENV:1088-1092
        }
        if (leg == 0) { obd[0] = e.cmdf[0]; obd[1] = e.cmdf[1]; obd[2] = e.cmdf[2]; }                        // ENV:1095-1097
        GaitCmd g = gait_cmd(P, e.cmdf);
        float t = cur_time(P, e);
        f3 toe;
        f3 last = e.jref;                                          // jointRefLast_ == previous jointRef_ (ENV:1887)
        if (flag_reset) last = leg_reference(P, g, leg, t - P.control_dt, toe);   // ENV:1799-1843
        f3 ref = leg_reference(P, g, leg, t, toe);
        const float inv_dt = 1.0f / P.control_dt;              // one (launch-uniform) division instead of three per lane
        e.jdref = mk((ref.x - last.x) * inv_dt, (ref.y - last.y) * inv_dt, (ref.z - last.z) * inv_dt);   // ENV:1886
        e.jref = ref;
        e.eeref = mk(toe.x + 0.19f * e.lm.sx, toe.y + 0.058f * e.lm.sy, toe.z);   // ENV:331-334, 1882-1889
    } else {
        const float* row = P.ref + (size_t)ref_row(P, e.frame_idx) * 30;       // ENV:1102-1106, 1670-1671
        e.cmdf[0] = row[27]; e.cmdf[1] = row[28]; e.cmdf[2] = row[29];
        if (leg == 0) { obd[0] = e.cmdf[0]; obd[1] = e.cmdf[1]; obd[2] = e.cmdf[2]; }
        e.jref = mk(row[3 * leg], row[3 * leg + 1], row[3 * leg + 2]);
        e.jdref = mk(row[12 + 3 * leg], row[12 + 3 * leg + 1], row[12 + 3 * leg + 2]);
    }
}

// time-based contact flag (ENV:1169-1188) or the detected one
__device__ __forceinline__ float contact_flag(const EnvParams& P, const EnvRegs& e, int leg, int detected) {
    if (!P.flag_time_contact) return detected ? 1.f : 0.f;
    return (gait_phase(P, e, leg) < P.lam) ? 1.f : 0.f;
}

// ------------------------------------------------------------------ reset (ENV:547-635)
static __device__ __noinline__ void reset_env(const EnvParams& P, const DevState& S, int r, int gid, int leg, uint32_t tick, EnvRegs& e) {
    e.itera += 1;                                                                  // ENV:554
    uint4 rr = philox(P.seed, gid, tick, P_RST_TIME_CMD);
    e.t0 = P.flag_manual ? 0.f : u01(rr.w);                                        // ENV:557
    e.frame_idx = 0;                                                               // ENV:565-568
    if (!P.flag_manual_traj && !P.flag_manual) {                                   // ENV:571 + sampling_reshape ENV:71-81
        double ratio = u01(rr.z);
        double rs = (ratio < 0.5 && ratio > 0) ? ratio * 4.0 / 3.0 : (2.0 * ratio + 1.0) / 3.0;
        e.frame_idx = int((P.frame_max - P.frame_len - 10) * rs);
        // the episode clock and the table row are separate counters in the reference (current_time_ ENV:557,631,786 vs frame_idx):
        // cur_time() = t0 + frame_idx * dt, so the random start row is taken out of t0 here
        e.t0 = (float)((double)e.t0 - (double)e.frame_idx * (double)P.control_dt);
    }
    e.cmdf[0] = e.cmdf[1] = e.cmdf[2] = 0.f;                                       // ENV:559-562
    e.torque_last = mk(0.f, 0.f, 0.f);                                             // ENV:575
    command_obs_update(P, S, r, gid, leg, tick, P_CMD + P_IN_RESET + 16, true, e); // ENV:577
    rr = philox(P.seed, gid, tick, P_RST_INIT);
    float n0 = usym(rr.x), n1 = usym(rr.y);                                        // shared scalar
    f3 qi = mk(e.jref.x * (n0 * 0.3f) + e.jref.x, e.jref.y * (n0 * 0.3f) + e.jref.y, e.jref.z * (n0 * 0.3f) + e.jref.z);          // ENV:584
    f3 qdi = mk(e.jdref.x * (n1 * 0.3f) + e.jdref.x, e.jdref.y * (n1 * 0.3f) + e.jdref.y, e.jdref.z * (n1 * 0.3f) + e.jdref.z);  // ENV:586
    rr = philox(P.seed, gid, tick, P_RST_BASEVEL);
    float vx = e.cmdf[0] * (usym(rr.x) * 0.2f + 1.0f); if (P.flag_wildcat) vx = -vx;                // ENV:588-589
    float vy = e.cmdf[1] * (usym(rr.y) * 0.2f + 1.0f);                                              // ENV:590
    float wz = e.cmdf[2] * (usym(rr.z) * 0.2f + 1.0f);                                              // ENV:591
    if (P.flag_manual) {                                                            // ENV:616-619
        e.b.p = mk(0.f, 0.f, 0.35f); e.b.v = mk(0, 0, 0); e.b.w = mk(0, 0, 0); e.q = nominal_q(P, leg); e.qd = mk(0, 0, 0);
    } else {                                                                        // ENV:593-606, 620-623
        rr = philox(P.seed, gid, tick, P_RST_XY);
        float a = u01(rr.x), bb = u01(rr.y);
        e.b.p = mk(a * 5.0f + (1.0f - a) * -5.0f, bb * 5.0f + (1.0f - bb) * -5.0f, 0.35f);
        e.b.v = mk(vx, vy, 0.f); e.b.w = mk(0.f, 0.f, wz); e.q = qi; e.qd = qdi;
        if (P.terrain) { float gh; f3 gn; terrain_sample(P, e.b.p.x, e.b.p.y, gh, gn); e.b.p.z += gh; }   // drop height measured from the terrain
    }
    e.b.qw = 1.f; e.b.qx = e.b.qy = e.b.qz = 0.f;
    e.contact_flag = 0.f; e.impulse_norm = 0.f;                                     // contact list empty after setState
    ObsOut o; update_observation(P, S, r, gid, leg, tick, P_IN_RESET, e, o);        // ENV:625
    if (P.flag_obs_filter) {                                                        // ENV:626 obDouble_last_ = obDouble_
        // every lane copies exactly the entries it wrote itself: no cross-lane ordering (reset runs divergent)
        const float* obd = S.obd + (size_t)r * 36; float* last = S.obd_last + (size_t)r * 36;
        for (int kk = 0; kk < 3; ++kk) { last[5 + 3 * leg + kk] = obd[5 + 3 * leg + kk]; last[17 + 3 * leg + kk] = obd[17 + 3 * leg + kk]; }
        if (leg == 0) { for (int i = 0; i < 5; ++i) last[i] = obd[i]; for (int i = 29; i < 35; ++i) last[i] = obd[i]; }
    }
    e.contact_flag = contact_flag(P, e, leg, 0);                                    // ENV:627
    command_obs_update(P, S, r, gid, leg, tick, P_CMD + P_IN_RESET, false, e);      // ENV:628
    e.frame_idx += 1;                                                               // ENV:630-631
    e.ep_len = 0; e.ep_ret = 0.f;
}

// scaled observation (ENV:1248-1262 without the ObsFilter branch): lane `leg` emits its share of the row
__device__ __forceinline__ void write_scaled_obs(const EnvParams& P, const DevState& S, int r, int leg, float* ob_row) {
    const float* obd = S.obd + (size_t)r * 36;
    f3 qn = nominal_q(P, leg);
    const float ivstd[3] = {1.0f / 5.f, 1.0f / 35.f, 1.0f / 40.f};   // reciprocal std: constant multiplies instead of IEEE divisions (<= 1 ulp)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ob_row[5 + 3 * leg + k] = (obd[5 + 3 * leg + k] - comp(qn, k)) / 1.0f;
        ob_row[17 + 3 * leg + k] = (obd[17 + 3 * leg + k] - 0.f) * ivstd[k];
    }
    if (leg == 0) {
        ob_row[0] = (obd[0] - (P.Vx_max + P.Vx_min) / 2.f) / 1.0f;                  // ENV:375, 383
        ob_row[1] = (obd[1] - (P.Vy_max + P.Vy_min) / 2.f) / 1.0f;
        ob_row[2] = (obd[2] - (P.omega_max + P.omega_min) / 2.f) / 1.0f;
        ob_row[3] = obd[3]; ob_row[4] = obd[4];
        ob_row[29] = obd[29] * (1.0f / 0.7f); ob_row[30] = obd[30] * (1.0f / 0.7f); ob_row[31] = (obd[31] - 1.0f) * (1.0f / 0.7f);
        ob_row[32] = obd[32] * (1.0f / 3.0f); ob_row[33] = obd[33] * (1.0f / 3.0f); ob_row[34] = obd[34] * (1.0f / 3.0f);
    }
}

// the same row from the values the step kernel still holds in registers (no read-back of obDouble_ through global memory); cmd = the
// obDouble_[0..2] entries command_obs_update wrote
__device__ __forceinline__ void write_scaled_obs_regs(const EnvParams& P, int leg, const ObsOut& o, const float* cmd, float* ob_row) {
    f3 qn = nominal_q(P, leg);
    const float ivstd[3] = {1.0f / 5.f, 1.0f / 35.f, 1.0f / 40.f};   // reciprocal std: constant multiplies instead of IEEE divisions (<= 1 ulp)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ob_row[5 + 3 * leg + k] = (o.oq[k] - comp(qn, k)) / 1.0f;
        ob_row[17 + 3 * leg + k] = (o.oqd[k] - 0.f) * ivstd[k];
    }
    if (leg == 0) {
        ob_row[0] = (cmd[0] - (P.Vx_max + P.Vx_min) / 2.f) / 1.0f;
        ob_row[1] = (cmd[1] - (P.Vy_max + P.Vy_min) / 2.f) / 1.0f;
        ob_row[2] = (cmd[2] - (P.omega_max + P.omega_min) / 2.f) / 1.0f;
        ob_row[3] = o.ph3; ob_row[4] = o.ph4;
        ob_row[29] = o.ob29 * (1.0f / 0.7f); ob_row[30] = o.ob30 * (1.0f / 0.7f); ob_row[31] = (o.ob31 - 1.0f) * (1.0f / 0.7f);
        ob_row[32] = o.om[0] * (1.0f / 3.0f); ob_row[33] = o.om[1] * (1.0f / 3.0f); ob_row[34] = o.om[2] * (1.0f / 3.0f);
    }
}

// ------------------------------------------------------------------ THE step kernel
// register cap of the step kernel: STEP_MINWARPS resident warps per SM -> 65536 / (32 * STEP_MINWARPS) registers per thread
#ifndef STEP_MINWARPS
#define STEP_MINWARPS 8
#endif
// BLK / SYNC: 64-thread CTAs without phase barriers up to ~5k robots (one warp per scheduler: pure latency), 128-thread CTAs with
// phase barriers above (launch_env_step)
template <int BLK, bool SYNC, bool TERR>
__global__ void __launch_bounds__(BLK, (STEP_MINWARPS * 32 + BLK - 1) / BLK) env_step_kernel(const __grid_constant__ StepArgs A) {
    const EnvParams& P = A.P; const DevState& S = A.S;
    const int tid = blockIdx.x * BLK + threadIdx.x;
    int r = A.r_begin + (tid >> 2); const int leg = tid & 3;
    const bool valid = r < A.r_end;
    if (!valid) r = A.r_end - 1;                  // tail lanes shadow the last robot (no stores) so quads stay convergent
    const int gid = r + (int)P.env_offset;
    EnvRegs e; load_env_hot(S, r, leg, e);

    // ---- action -> joint targets (ENV:700-707)
    float an = 0.f;
    if (P.action_noise != 0.f) an = usym(philox(P.seed, gid, A.tick, P_ACT).x);
    f3 qn = nominal_q(P, leg);
    const float* act = A.action + (size_t)r * ACT_DIM + 3 * leg;
    f3 pt;
    {
        float p0 = act[0] * 1.0f + qn.x, p1 = act[1] * 1.0f + qn.y, p2 = act[2] * 1.0f + qn.z;
        p0 = (1.0f - P.filter_para) * p0 + P.filter_para * e.ptl.x;
        p1 = (1.0f - P.filter_para) * p1 + P.filter_para * e.ptl.y;
        p2 = (1.0f - P.filter_para) * p2 + P.filter_para * e.ptl.z;
        float k = P.action_noise * an;             // one scalar for all joints
        pt = mk(p0 * k + p0, p1 * k + p1, p2 * k + p2);
    }
    e.ptl = pt;
    // ---- ForceDisturbance + Manual: state_disturbance every 10 gait periods (ENV:744-747, 912-940); without Manual the
    // reference's force_attack never fires
    if (P.flag_force_dist && P.flag_manual) {
        if (P.disturb_every > 0 && S.frame_idx[r] % P.disturb_every == 0) {
            const uint4 ra = philox(P.seed, gid, A.tick, P_DISTURB), rb = philox(P.seed, gid, A.tick, P_DISTURB + 1);
            const float ratio = 0.5f;
            e.b.p.z += 0.03f * usym(ra.x) * ratio;
            e.b.qw += 0.1f * usym(ra.y) * ratio; e.b.qx += 0.1f * usym(ra.z) * ratio; e.b.qy += 0.1f * usym(ra.w) * ratio; e.b.qz += 0.1f * usym(rb.x) * ratio;
            const float qn = 1.0f / sqrtf(e.b.qw * e.b.qw + e.b.qx * e.b.qx + e.b.qy * e.b.qy + e.b.qz * e.b.qz);
            e.b.qw *= qn; e.b.qx *= qn; e.b.qy *= qn; e.b.qz *= qn;
            e.b.v.z += 0.1f * usym(rb.y) * ratio; e.b.w.x += 0.3f * usym(rb.z) * ratio; e.b.w.y += 0.3f * usym(rb.w) * ratio;
        }
    }
    // ---- PD + torque filter + clamp + integrate, loop_count times (ENV:758-774)
    const float kp0 = P.stiffness * P.abad_ratio, kd0 = P.damping * P.abad_ratio;
    const float rr_ = P.motor_max_torque / (P.motor_max_speed - P.motor_crit_speed);
    const float ilow_ = 1.0f / (-P.motor_max_speed + P.motor_crit_speed);      // hoisted out of the substep loop
    f3 tau = e.torque_last;      // a step with zero substeps (simulation_dt > control_dt, int() ENV:711) sees the stale member `torque`, which equals torque_last between steps (ENV:1511-1515)
    ContactOut co; co.foot_active = 0; co.foot_impulse = mk(0, 0, 0); co.sweeps = 0;
    const float box_reach = trunk_box_reach(P);
    // PD + torque filter + torque_clamp (ENV:761-767, 1273-1305) on the current joint state
    auto pd_torque = [&]() {
        float t0 = (pt.x - e.q.x) * kp0 - e.qd.x * kd0;
        float t1 = (pt.y - e.q.y) * P.stiffness - e.qd.y * P.damping;
        float t2 = (pt.z - e.q.z) * P.stiffness - e.qd.z * P.damping;
        t0 = 0.99f * t0 + (1.0f - 0.99f) * e.torque_last.x;
        t1 = 0.99f * t1 + (1.0f - 0.99f) * e.torque_last.y;
        t2 = 0.99f * t2 + (1.0f - 0.99f) * e.torque_last.z;
        float tt[3] = {t0, t1, t2}, qv[3] = {e.qd.x, e.qd.y, e.qd.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float ratio = (k == 2) ? 1.55f : 1.0f;
            float s = qv[k] * ratio;
            float up = (s > P.motor_crit_speed) ? (P.motor_max_torque - (s - P.motor_crit_speed) * rr_) : P.motor_max_torque;
            up = up * ratio;
            float low = (s < -P.motor_crit_speed) ? ((-P.motor_max_speed - s) * ilow_ * -P.motor_max_torque) : -P.motor_max_torque;
            low = low * ratio;
            tt[k] = fmaxf(fminf(tt[k], up), low);
        }
        tau = mk(tt[0], tt[1], tt[2]);
    };
    // hot loop: substeps without the trunk-box contact code; it hands over (state untouched) at the first substep in which a box corner
    // touches the ground in this warp / CTA -- a fallen robot, at most the last control step of an episode -- and the complete version
    // finishes the control step
    int it = 0;
#pragma unroll 1
    for (; it < P.loop_count; ++it) {
        pd_torque();
        if (!integrate_substep<SYNC, TERR, false>(P, e.lm, e.bm, leg, e.b, e.q, e.qd, tau, co, box_reach)) break;
    }
#pragma unroll 1
    for (; it < P.loop_count; ++it) {
        pd_torque();
        integrate_substep<SYNC, TERR, true>(P, e.lm, e.bm, leg, e.b, e.q, e.qd, tau, co, box_reach);
    }
    load_env_cold(S, r, leg, e);                   // clock, command, references, counters: first needed here
    e.tau_applied = tau;

    // ---- observation (ENV:776)
    ObsOut o; update_observation(P, S, r, gid, leg, A.tick, 0, e, o);
    // ---- contact information (ENV:777, 1199-1231): |impulse| / control_dt, toe-frame speed
    f3 bx, by, bz; quat_cols(e.b.qw, e.b.qx, e.b.qy, e.b.qz, bx, by, bz);
    LegKin k; leg_fk(P, e.lm, bx, by, bz, e.q, k);
    f3 w1 = axpy(e.qd.x, k.a1, e.b.w), w2 = axpy(e.qd.y, k.a2, w1), w3 = axpy(e.qd.z, k.a2, w2);
    f3 vtoe = e.b.v + cross(e.b.w, k.j1) + cross(w1, k.j2 - k.j1) + cross(w2, k.j3 - k.j2) + cross(w3, k.toe - k.j3);
    float vel_norm = sqrtf(dot(vtoe, vtoe));
    float imp = co.foot_active ? sqrtf(dot(co.foot_impulse, co.foot_impulse)) : 0.f;
    const float inv_dt = 1.0f / P.control_dt;
    float force_norm = imp * inv_dt;              // ENV:1208 (control_dt, quirk 4)
    e.impulse_norm = imp;

    float ground_z = 0.f;
    if (P.terrain) { f3 gn; terrain_sample(P, e.b.p.x, e.b.p.y, ground_z, gn); }
    // ---- reward (ENV:1444-1548)
    float rew, r_ee, r_pos, r_att, r_joint, r_vel;
    {
        f3 ee = mk(dot(bx, k.toe), dot(by, k.toe), dot(bz, k.toe));       // R^T (p_toe - p_base)  ENV:1452-1456
        f3 de = ee - e.eeref;
        float ees = qsum(dot(de, de));
        r_ee = P.ee_coeff * expf(-40.f * ees);
        float dz = e.b.p.z - ground_z - P.stand_height;              // height above the terrain under the trunk (flat ground: z)
        r_pos = P.pos_coeff * expf(-80.f * (dz * dz));
        r_att = P.atti_coeff * expf(-80.f * (o.ob29 * o.ob29 + o.ob30 * o.ob30));
        f3 dj = e.jref - e.q, dd = e.jdref - e.qd;
        float jr = qsum(dot(dj, dj)), jd = qsum(dot(dd, dd));
        r_joint = P.joint_coeff * 0.25f * expf(-2.0f * jr);
        float r_jd = P.joint_coeff * 0.75f * expf(-P.control_dt * jd);
        f3 lref = mk(P.flag_wildcat ? -e.cmdf[0] : e.cmdf[0], e.cmdf[1], 0.f), aref = mk(0.f, 0.f, e.cmdf[2]);
        f3 le = o.blin - lref, ae = o.bang - aref;
        r_vel = P.vel_coeff / 2.f * expf(-2.f * dot(le, le)) + P.vel_coeff / 2.f * expf(-2.f * dot(ae, ae));
        f3 tn = mk(tau.x * (1.0f / 18.f), tau.y * (1.0f / 18.f), tau.z * (1.0f / 27.f));            // ENV:354, 1511
        f3 dtq = tn - e.torque_last;
        float tns = qsum(dot(tn, tn)), tds = qsum(dot(dtq, dtq));
        float r_tq = P.torque_coeff / 2.0f * expf(-0.1f * tns) + P.torque_coeff / 2.0f * expf(-0.1f * inv_dt * tds);
        e.torque_last = tn;                                               // ENV:1515
        const float rp = gait_phase(P, e, leg);                           // ENV:1523-1527
        const float sraw = smooth_raw(rp, 2.f, P.lam);                    // both shaping functions clamp the same raw value (ENV:118-156)
        float cr = 4.f * vel_norm * vel_norm * smooth_clamp(sraw) +
                   2.f * (force_norm * (1.0f / 12.5f)) * (force_norm * (1.0f / 12.5f)) * smooth_clamp2(sraw);
        cr = qsum(cr);
        float r_ct = P.contact_coeff * expf(-2.f * cr);
        rew = (r_ee + r_pos + r_joint + r_jd + r_vel + r_att + r_tq + r_ct);   // ENV:1546-1547
    }
    // ---- next reference (ENV:784-787)
    command_obs_update(P, S, r, gid, leg, A.tick, P_CMD, false, e);
    e.contact_flag = contact_flag(P, e, leg, co.foot_active);
    e.frame_idx += 1;
    // ---- VEC::perAgentStep (VEC:352-372)
    const float zrel = e.b.p.z - ground_z;
    // ENV:1560 (relative to the terrain under the trunk); written so that a non-finite state also terminates and resets the env
    const bool done = !((zrel >= 0.15f) && (zrel <= 0.65f) && (o.ob31 >= 0.5f));
    const float base_z = e.b.p.z;
    e.ep_len += 1;
    float ep_ret_out = 0.f; int ep_len_out = 0;
    if (done) {
        rew += P.terminal_coeff;                                          // VEC:370
        e.ep_ret += rew; ep_ret_out = e.ep_ret; ep_len_out = e.ep_len;
        EnvRegs tmp = e;                                                  // only the copy has its address taken (rare path)
        reset_env(P, S, r, gid, leg, A.tick, tmp);                        // VEC:369
        e = tmp;
    } else {
        e.ep_ret += rew;
    }
    if (valid) {
        store_env(S, r, leg, e);
        if (leg == 0) {
            A.reward[r] = rew; A.done[r] = done ? 1 : 0;
            if (A.extra) {
                float* ex = A.extra + (size_t)r * EXTRA_DIM;              // ENV:944-949 insertion order
                ex[0] = r_ee; ex[1] = r_pos; ex[2] = base_z; ex[3] = r_att; ex[4] = r_joint; ex[5] = r_vel;
            }
            if (A.ep_ret_out) { A.ep_ret_out[r] = ep_ret_out; A.ep_len_out[r] = ep_len_out; }
            S.solver_sweeps[r] = co.sweeps;
        }
    }
    // every lane scales exactly the obDouble_ entries it wrote itself (no cross-lane dependency): from registers, except after a reset
    // (reset_env rewrote obDouble_) and in Manual mode (the command entries are written by SetCommand), where they are read back
    if (!P.flag_obs_filter && A.ob && valid) {
        if (done || P.flag_manual) write_scaled_obs(P, S, r, leg, A.ob + (size_t)r * OB_DIM);
        else write_scaled_obs_regs(P, leg, o, e.cmdf, A.ob + (size_t)r * OB_DIM);
    }
}

__global__ void __launch_bounds__(BLOCK) env_reset_kernel(const __grid_constant__ StepArgs A) {
    const EnvParams& P = A.P; const DevState& S = A.S;
    const int tid = blockIdx.x * BLOCK + threadIdx.x;
    int r = tid >> 2; const int leg = tid & 3;
    const bool valid = r < P.N; if (!valid) r = P.N - 1;
    const int gid = r + (int)P.env_offset;
    EnvRegs e; load_env(S, r, leg, e);
    reset_env(P, S, r, gid, leg, A.tick, e);
    if (valid) store_env(S, r, leg, e);
    if (!P.flag_obs_filter && A.ob && valid) write_scaled_obs(P, S, r, leg, A.ob + (size_t)r * OB_DIM);
}

// observe (ENV:1248-1262) incl. the stateful ObsFilter branch; one thread per env
__global__ void env_observe_kernel(EnvParams P, DevState S, float* ob) {
    int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= P.N) return;
    float* obd = S.obd + (size_t)r * 36;
    if (P.flag_obs_filter) {
        float* last = S.obd_last + (size_t)r * 36;
        for (int i = 5; i < 35; ++i) obd[i] = obd[i] * P.obs_filter_alpha + last[i] * (1.0f - P.obs_filter_alpha);
        for (int i = 0; i < 35; ++i) last[i] = obd[i];
    }
    float* o = ob + (size_t)r * OB_DIM;
    const float ivstd[3] = {1.0f / 5.f, 1.0f / 35.f, 1.0f / 40.f};   // reciprocal std: constant multiplies instead of IEEE divisions (<= 1 ulp)
    o[0] = (obd[0] - (P.Vx_max + P.Vx_min) / 2.f) / 1.0f; o[1] = (obd[1] - (P.Vy_max + P.Vy_min) / 2.f) / 1.0f;
    o[2] = (obd[2] - (P.omega_max + P.omega_min) / 2.f) / 1.0f; o[3] = obd[3]; o[4] = obd[4];
    for (int j = 0; j < 12; ++j) {
        float qn = (j % 3 == 0) ? (((j / 3) & 1) ? P.abad : -P.abad) : ((j % 3 == 1) ? -0.78f : 1.57f);
        o[5 + j] = (obd[5 + j] - qn) / 1.0f; o[17 + j] = obd[17 + j] * ivstd[j % 3];
    }
    o[29] = obd[29] * (1.0f / 0.7f); o[30] = obd[30] * (1.0f / 0.7f); o[31] = (obd[31] - 1.0f) * (1.0f / 0.7f);
    o[32] = obd[32] * (1.0f / 3.0f); o[33] = obd[33] * (1.0f / 3.0f); o[34] = obd[34] * (1.0f / 3.0f);
}

// ------------------------------------------------------------------ probes: M (row-major 18x18), M^-1 (ENV:1375-1391), h (ENV:1396-1402)
__global__ void __launch_bounds__(BLOCK) env_probe_kernel(EnvParams P, DevState S, float* Mout, float* Minv, float* hout) {
    const int tid = blockIdx.x * BLOCK + threadIdx.x;
    int r = tid >> 2; const int leg = tid & 3;
    const bool valid = r < P.N; if (!valid) r = P.N - 1;
    EnvRegs e; load_env(S, r, leg, e);
    f3 bx, by, bz; quat_cols(e.b.qw, e.b.qx, e.b.qy, e.b.qz, bx, by, bz);
    LegKin k; leg_fk(P, e.lm, bx, by, bz, e.q, k);
    Dyn d; float A[21]; S3 D;
    dynamics(P, e.lm, e.bm, e.b, bx, by, bz, k, e.qd, d, A, &D);
    if (valid && hout) {
        float* h = hout + (size_t)r * 18;
        if (leg == 0) for (int a = 0; a < 6; ++a) h[a] = d.hbv(a);
        h[6 + 3 * leg] = d.hl.x; h[7 + 3 * leg] = d.hl.y; h[8 + 3 * leg] = d.hl.z;
    }
    if (valid && Mout) {
        float* M = Mout + (size_t)r * 324;
        if (leg == 0) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) M[i * 18 + j] = A[i >= j ? tri(i, j) : tri(j, i)];
        int c0 = 6 + 3 * leg;
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c) { M[a * 18 + c0 + c] = d.Bv(a, c); M[(c0 + c) * 18 + a] = d.Bv(a, c); }
        for (int i = 0; i < 12; ++i) for (int c = 0; c < 3; ++c) if (i / 3 != leg) M[(c0 + c) * 18 + 6 + i] = 0.f;
        M[(c0 + 0) * 18 + c0 + 0] = D.xx; M[(c0 + 0) * 18 + c0 + 1] = D.xy; M[(c0 + 0) * 18 + c0 + 2] = D.xz;
        M[(c0 + 1) * 18 + c0 + 0] = D.xy; M[(c0 + 1) * 18 + c0 + 1] = D.yy; M[(c0 + 1) * 18 + c0 + 2] = D.yz;
        M[(c0 + 2) * 18 + c0 + 0] = D.xz; M[(c0 + 2) * 18 + c0 + 1] = D.yz; M[(c0 + 2) * 18 + c0 + 2] = D.zz;
    }
    if (Minv) {
        for (int col = 0; col < 18; ++col) {       // every lane takes part in the quad solve
            float rb[6]; for (int a = 0; a < 6; ++a) rb[a] = (a == col) ? 1.f : 0.f;
            int lc = col - 6 - 3 * leg;
            f3 rl = mk(lc == 0 ? 1.f : 0.f, lc == 1 ? 1.f : 0.f, lc == 2 ? 1.f : 0.f);
            float xb[6]; f3 xl; solve_full(d, rb, rl, xb, xl);
            if (valid) {
                float* mi = Minv + (size_t)r * 324 + (size_t)col * 18;      // column-major flatten == symmetric
                if (leg == 0) for (int a = 0; a < 6; ++a) mi[a] = xb[a];
                mi[6 + 3 * leg] = xl.x; mi[7 + 3 * leg] = xl.y; mi[8 + 3 * leg] = xl.z;
            }
        }
    }
}

// ------------------------------------------------------------------ Crutial: True -- meteor sphere (ENV:717-741, 815-861)
// Same specification as the oracle (oracle/bp5_oracle.hpp, meteor_update): the CubeNum coincident steel spheres of the reference are
// one sphere of CubeNum times the mass; sphere <-> trunk box (frictionless, restitution 0.95 above 1 mm/s) and sphere <-> ground are
// resolved once per control step, before the physics substeps, against the trunk pose at the start of the step.  The impulse on the
// robot goes through the same factorised mass matrix as the contact solver: du = -lambda M^-1 J^T n.
// respawn_only: called after irrl_reset so that GetSphereInfo sees the sphere the reset placed (ENV:608-611).
__global__ void __launch_bounds__(BLOCK) env_meteor_kernel(const __grid_constant__ StepArgs A, int respawn_only) {
    const EnvParams& P = A.P; const DevState& S = A.S;
    const int tid = blockIdx.x * BLOCK + threadIdx.x;
    int r = tid >> 2; const int leg = tid & 3;
    const bool valid = r < P.N; if (!valid) r = P.N - 1;
    EnvRegs e; load_env(S, r, leg, e);
    float* mp = S.meteor + (size_t)r * 12;
    float4 m0 = ld4(mp), m1 = ld4(mp + 4); const float m8 = mp[8];
    f3 sp = mk(m0.x, m0.y, m0.z), sv = mk(m0.w, m1.x, m1.y); int mode = (int)m1.z; float rad = m1.w, mass = m8;
    const bool periodic = P.meteor_every > 0 && (e.frame_idx % P.meteor_every) == 0;
    if (respawn_only || periodic || e.frame_idx == 1) {            // ENV:827-838 (frame_idx == 1: the env was reset since its last step)
        const float t = cur_time(P, e);
        rad = (t / 5.0f + 1.0f) * 0.08f; mass = (float)P.num_cube * (t / 5.0f + 0.2f);
        sp = mk(e.b.p.x + 0.05f, e.b.p.y, e.b.p.z + 1.0f); sv = mk(0.f, 0.f, 0.f); mode = 0;
    }
    if (!(respawn_only || periodic)) {
        if (mode == 0) { mode = 1; sv = mk(e.b.v.x, e.b.v.y, -5.0f); }             // ENV:849-858
        f3 bx, by, bz; quat_cols(e.b.qw, e.b.qx, e.b.qy, e.b.qz, bx, by, bz);
        LegKin k; leg_fk(P, e.lm, bx, by, bz, e.q, k);
        Dyn d; dynamics(P, e.lm, e.bm, e.b, bx, by, bz, k, e.qd, d, nullptr, nullptr, false, leg);
        float du[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; f3 dq = mk(0.f, 0.f, 0.f);
        const float dt = P.sim_dt;
        for (int sub = 0; sub < P.loop_count; ++sub) {
            sv.z -= P.gravity * dt;
            f3 dw = sp - e.b.p;
            f3 dl = mk(dot(bx, dw), dot(by, dw), dot(bz, dw));
            f3 q = mk(fminf(fmaxf(dl.x, -P.box_half[0]), P.box_half[0]), fminf(fmaxf(dl.y, -P.box_half[1]), P.box_half[1]), fminf(fmaxf(dl.z, -P.box_half[2]), P.box_half[2]));
            f3 del = dl - q; const float dist2 = dot(del, del);
            if (dist2 < rad * rad && dist2 > 1e-12f) {
                const float inv = rsqrtf(dist2);
                f3 nb = inv * del;
                f3 n = axpy(nb.x, bx, axpy(nb.y, by, nb.z * bz)), x = axpy(q.x, bx, axpy(q.y, by, q.z * bz));
                f3 vb = e.b.v + mk(du[0], du[1], du[2]), wb = e.b.w + mk(du[3], du[4], du[5]);
                f3 vpt = vb + cross(wb, x);
                const float vrel = dot(n, sv - vpt);
                if (vrel < 0.f) {
                    f3 xn = cross(x, n);
                    float z[6] = {n.x, n.y, n.z, xn.x, xn.y, xn.z};                  // J^T n restricted to the trunk coordinates
                    fwd6(d, z);
                    float G = 0.f;
#pragma unroll
                    for (int a = 0; a < 6; ++a) G = fmaf(z[a], z[a], G);
                    bwd6(d, z);                                                       // z = S^-1 J^T n  (trunk part of M^-1 J^T n)
                    const float er = (-vrel > 0.001f) ? 0.95f : 0.f;                  // steel-steel pair ENV:244
                    const float lam = -(1.0f + er) * vrel / (G + 1.0f / mass);
                    sv = axpy(lam / mass, n, sv);
#pragma unroll
                    for (int a = 0; a < 6; ++a) du[a] -= lam * z[a];
                    // joint part of M^-1 J^T n:  -Y^T z  (this lane's leg)
                    dq.x += lam * (d.Yv(0, 0) * z[0] + d.Yv(1, 0) * z[1] + d.Yv(2, 0) * z[2] + d.Yv(3, 0) * z[3] + d.Yv(4, 0) * z[4] + d.Yv(5, 0) * z[5]);
                    dq.y += lam * (d.Yv(0, 1) * z[0] + d.Yv(1, 1) * z[1] + d.Yv(2, 1) * z[2] + d.Yv(3, 1) * z[3] + d.Yv(4, 1) * z[4] + d.Yv(5, 1) * z[5]);
                    dq.z += lam * (d.Yv(0, 2) * z[0] + d.Yv(1, 2) * z[1] + d.Yv(2, 2) * z[2] + d.Yv(3, 2) * z[3] + d.Yv(4, 2) * z[4] + d.Yv(5, 2) * z[5]);
                }
            }
            float hh = 0.f; f3 nn = mk(0.f, 0.f, 1.f);
            if (P.terrain) terrain_sample(P, sp.x, sp.y, hh, nn);
            if ((sp.z - hh) * nn.z - rad <= 0.f) {
                const float vn = dot(nn, sv);
                if (vn < 0.f) {
                    sv = axpy(-vn, nn, sv);
                    const float vt = sqrtf(dot(sv, sv));
                    if (vt > 1e-9f) { const float dv = fminf(vt, 0.8f * (-vn)) / vt; sv = axpy(-dv, sv, sv); }
                }
            }
            sp = axpy(dt, sv, sp);
        }
        e.b.v = e.b.v + mk(du[0], du[1], du[2]); e.b.w = e.b.w + mk(du[3], du[4], du[5]);
        e.qd = e.qd + dq;
        if (valid) store_env(S, r, leg, e);
    }
    if (valid && leg == 0) { st4(mp, sp.x, sp.y, sp.z, sv.x); st4(mp + 4, sv.y, sv.z, (float)mode, rad); mp[8] = mass; }
}

// single physics substep with externally supplied joint torques (parity tests of world.integrate())
__global__ void __launch_bounds__(BLOCK) env_integrate_kernel(EnvParams P, DevState S, const float* tau12, float* contact_out) {
    const int tid = blockIdx.x * BLOCK + threadIdx.x;
    int r = tid >> 2; const int leg = tid & 3;
    const bool valid = r < P.N; if (!valid) r = P.N - 1;
    EnvRegs e; load_env(S, r, leg, e);
    const float* t = tau12 + (size_t)r * 12 + 3 * leg;
    f3 tau = mk(t[0], t[1], t[2]);
    ContactOut co; integrate_substep(P, e.lm, e.bm, leg, e.b, e.q, e.qd, tau, co, trunk_box_reach(P));
    e.tau_applied = tau; e.contact_flag = co.foot_active ? 1.f : 0.f;
    e.impulse_norm = co.foot_active ? sqrtf(dot(co.foot_impulse, co.foot_impulse)) : 0.f;
    if (valid) {
        store_env(S, r, leg, e);
        if (contact_out) {   // [N][4][4]: active, impulse xyz
            float* c = contact_out + ((size_t)r * 4 + leg) * 4;
            c[0] = (float)co.foot_active; c[1] = co.foot_impulse.x; c[2] = co.foot_impulse.y; c[3] = co.foot_impulse.z;
        }
        if (leg == 0) S.solver_sweeps[r] = co.sweeps;
    }
}

// ------------------------------------------------------------------ flat state vector <-> SoA (layout in include/irrl_b200.h)
__global__ void env_get_state_kernel(EnvParams P, DevState S, float* out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= P.N) return;
    float* s = out + (size_t)r * STATE_DIM;
    const float* b = S.base + (size_t)r * 16;
    for (int i = 0; i < 7; ++i) s[i] = b[i];
    for (int i = 0; i < 6; ++i) s[19 + i] = b[7 + i];
    for (int l = 0; l < 4; ++l) {
        const float* lp = S.legs + ((size_t)r * 4 + l) * 16; const float* l2 = S.legs2 + ((size_t)r * 4 + l) * 12;
        for (int k = 0; k < 3; ++k) {
            s[7 + 3 * l + k] = lp[k]; s[25 + 3 * l + k] = lp[3 + k]; s[49 + 3 * l + k] = lp[6 + k];
            s[67 + 3 * l + k] = lp[9 + k]; s[79 + 3 * l + k] = lp[12 + k];
            s[37 + 3 * l + k] = l2[k]; s[91 + 3 * l + k] = l2[3 + k]; s[179 + 3 * l + k] = l2[6 + k];
        }
        s[105 + l] = l2[9];
    }
    const float* c = S.cmd + (size_t)r * 8;
    for (int i = 0; i < 3; ++i) { s[61 + i] = c[i]; s[64 + i] = c[3 + i]; }
    s[103] = b[13]; s[104] = (float)S.frame_idx[r];
    for (int i = 0; i < 35; ++i) { s[109 + i] = S.obd[(size_t)r * 36 + i]; s[144 + i] = S.obd_last[(size_t)r * 36 + i]; }
    s[191] = (float)S.itera[r];
}
__global__ void env_set_state_kernel(EnvParams P, DevState S, const float* in) {
    int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= P.N) return;
    const float* s = in + (size_t)r * STATE_DIM;
    float* b = S.base + (size_t)r * 16;
    for (int i = 0; i < 7; ++i) b[i] = s[i];
    for (int i = 0; i < 6; ++i) b[7 + i] = s[19 + i];
    b[13] = s[103]; b[14] = 0.f; b[15] = 0.f;
    for (int l = 0; l < 4; ++l) {
        float* lp = S.legs + ((size_t)r * 4 + l) * 16; float* l2 = S.legs2 + ((size_t)r * 4 + l) * 12;
        for (int k = 0; k < 3; ++k) {
            lp[k] = s[7 + 3 * l + k]; lp[3 + k] = s[25 + 3 * l + k]; lp[6 + k] = s[49 + 3 * l + k];
            lp[9 + k] = s[67 + 3 * l + k]; lp[12 + k] = s[79 + 3 * l + k];
            l2[k] = s[37 + 3 * l + k]; l2[3 + k] = s[91 + 3 * l + k]; l2[6 + k] = s[179 + 3 * l + k];
        }
        lp[15] = 0.f; l2[9] = s[105 + l]; l2[10] = 0.f; l2[11] = 0.f;
    }
    float* c = S.cmd + (size_t)r * 8;
    for (int i = 0; i < 3; ++i) { c[i] = s[61 + i]; c[3 + i] = s[64 + i]; }
    c[6] = c[7] = 0.f;
    S.frame_idx[r] = (int)s[104];
    for (int i = 0; i < 35; ++i) { S.obd[(size_t)r * 36 + i] = s[109 + i]; S.obd_last[(size_t)r * 36 + i] = s[144 + i]; }
    S.itera[r] = (int)s[191];
}

// ------------------------------------------------------------------ init: nominal / domain-randomised model (ENV:435-477)
__global__ void env_init_kernel(EnvParams P, DevState S) {
    int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= P.N) return;
    const uint32_t gid = r + P.env_offset, T = 0xFFFFFFFFu;
    float mu = P.mu, rest = P.restitution, thr = P.rest_threshold;
    float mass[13], com[13][3];
    mass[0] = P.m0; com[0][0] = P.com0[0]; com[0][1] = P.com0[1]; com[0][2] = P.com0[2];
    for (int l = 0; l < 4; ++l) {
        float sx = (l < 2) ? 1.f : -1.f, sy = (l & 1) ? 1.f : -1.f;
        mass[1 + 3 * l] = P.m1; com[1 + 3 * l][0] = P.com1[0] * sx; com[1 + 3 * l][1] = P.com1[1] * sy; com[1 + 3 * l][2] = P.com1[2];
        mass[2 + 3 * l] = P.m2; com[2 + 3 * l][0] = P.com2[0];      com[2 + 3 * l][1] = P.com2[1] * sy; com[2 + 3 * l][2] = P.com2[2];
        mass[3 + 3 * l] = P.m3; com[3 + 3 * l][0] = 0.f;            com[3 + 3 * l][1] = 0.f;            com[3 + 3 * l][2] = P.com3z;
    }
    float knee_z = P.knee_z;
    if (P.flag_stochastic) {
        uint4 rr = philox(P.seed, gid, T, P_DR_MATERIAL);
        mu = u01(rr.x) * 0.6f + 0.4f; rest = u01(rr.y) * 0.3f; thr = u01(rr.z) * 2.0f;                 // ENV:440-442
        for (int i = 0; i < 13; ++i) {
            rr = philox(P.seed, gid, T, P_DR_MASS + i);
            mass[i] = mass[i] * ((u01(rr.x) - 0.5f) / 0.5f * 0.15f + 1.0f);                            // ENV:454-456, 2069
            rr = philox(P.seed, gid, T, P_DR_COM + i);
            com[i][0] += usym(rr.x) * 0.02f; com[i][1] += usym(rr.y) * 0.02f; com[i][2] += usym(rr.z) * 0.02f;   // ENV:463-465, 2070
        }
        rr = philox(P.seed, gid, T, P_DR_CALF);
        knee_z += (u01(rr.x) - 0.5f) / 0.5f * 0.01f;                                                   // ENV:472-476, 2071
    }
    float* bm = S.basemodel + (size_t)r * 8;
    bm[0] = mass[0]; bm[1] = com[0][0]; bm[2] = com[0][1]; bm[3] = com[0][2]; bm[4] = mu; bm[5] = rest; bm[6] = thr; bm[7] = 0.f;
    for (int l = 0; l < 4; ++l) {
        float* m = S.legmodel + ((size_t)r * 4 + l) * 16;
        m[0] = mass[1 + 3 * l]; m[1] = mass[2 + 3 * l]; m[2] = mass[3 + 3 * l]; m[3] = knee_z;
        for (int b = 0; b < 3; ++b) { for (int a = 0; a < 3; ++a) m[4 + 4 * b + a] = com[1 + 3 * l + b][a]; m[4 + 4 * b + 3] = 0.f; }
        // initial state = gc_init_ (ENV:317-322), everything else zero
        float* lp = S.legs + ((size_t)r * 4 + l) * 16; float* l2 = S.legs2 + ((size_t)r * 4 + l) * 12;
        for (int i = 0; i < 16; ++i) lp[i] = 0.f;
        for (int i = 0; i < 12; ++i) l2[i] = 0.f;
        lp[0] = (l & 1) ? P.abad : -P.abad; lp[1] = -0.78f; lp[2] = 1.57f;
        lp[9] = (l & 1) ? P.abad : -P.abad;                                                             // jointRef_ ENV:415-418
    }
    float* b = S.base + (size_t)r * 16; for (int i = 0; i < 16; ++i) b[i] = 0.f; b[2] = 0.35f; b[3] = 1.f;
    float* c = S.cmd + (size_t)r * 8; for (int i = 0; i < 8; ++i) c[i] = 0.f;
    for (int i = 0; i < 36; ++i) { S.obd[(size_t)r * 36 + i] = 0.f; S.obd_last[(size_t)r * 36 + i] = 0.f; }
    S.frame_idx[r] = 0; S.itera[r] = 0; S.ep_len[r] = 0; S.ep_ret[r] = 0.f; S.solver_sweeps[r] = 0;
}

// ------------------------------------------------------------------ host launchers
static inline int quad_grid(int N) { return (N * 4 + BLOCK - 1) / BLOCK; }
void launch_env_step(const StepArgs& a, cudaStream_t st) {
    static const int sync_min = [] { const char* e = getenv("IRRL_STEP_SYNC_MIN"); return e ? atoi(e) : 5120; }();   // tuning knob: robots above which the barrier variant runs (measured: 4096 -> 67.8 vs 73.6 us without / with barriers, 6144 -> 79.4 vs 77.5)
    static const int blk = [] { const char* e = getenv("IRRL_STEP_BLK"); return e ? atoi(e) : 0; }();   // experiment: force the CTA size (256 = one lock-stepped CTA per SM)
    const int n = a.r_end - a.r_begin; if (n <= 0) return;
    const bool terr = a.P.terrain != nullptr;   // launch-uniform: the flat-ground kernels carry no heightfield code in the substep loop
#define STEP_LAUNCH(B, SY) do { if (terr) env_step_kernel<B, SY, true><<<(n * 4 + B - 1) / B, B, 0, st>>>(a); \
                               else env_step_kernel<B, SY, false><<<(n * 4 + B - 1) / B, B, 0, st>>>(a); } while (0)
    if (blk == 256) STEP_LAUNCH(256, true);
    else if (blk == 128 || (blk == 0 && n > sync_min)) STEP_LAUNCH(128, true);
    else STEP_LAUNCH(64, false);
#undef STEP_LAUNCH
}
void launch_env_meteor(const StepArgs& a, int respawn_only, cudaStream_t st) { env_meteor_kernel<<<quad_grid(a.P.N), BLOCK, 0, st>>>(a, respawn_only); }
void launch_env_reset(const StepArgs& a, cudaStream_t st) { env_reset_kernel<<<quad_grid(a.P.N), BLOCK, 0, st>>>(a); }
void launch_env_observe(const EnvParams& P, const DevState& S, float* ob, cudaStream_t st) { env_observe_kernel<<<(P.N + 127) / 128, 128, 0, st>>>(P, S, ob); }
void launch_env_probe(const EnvParams& P, const DevState& S, float* M, float* Minv, float* h, cudaStream_t st) { env_probe_kernel<<<quad_grid(P.N), BLOCK, 0, st>>>(P, S, M, Minv, h); }
void launch_env_integrate(const EnvParams& P, const DevState& S, const float* tau, float* contact_out, cudaStream_t st) { env_integrate_kernel<<<quad_grid(P.N), BLOCK, 0, st>>>(P, S, tau, contact_out); }
void launch_env_get_state(const EnvParams& P, const DevState& S, float* out, cudaStream_t st) { env_get_state_kernel<<<(P.N + 127) / 128, 128, 0, st>>>(P, S, out); }
void launch_env_set_state(const EnvParams& P, const DevState& S, const float* in, cudaStream_t st) { env_set_state_kernel<<<(P.N + 127) / 128, 128, 0, st>>>(P, S, in); }
void launch_env_init(const EnvParams& P, const DevState& S, cudaStream_t st) { env_init_kernel<<<(P.N + 127) / 128, 128, 0, st>>>(P, S); }

}  // namespace irrl
