// Device-side building blocks of the bp5 environment step for sm_100a.
//
// Mapping: ONE ROBOT = ONE QUAD OF LANES (lane l of the quad owns leg l: FR, FL, HR, HL); a warp carries
// 8 robots.  The kinematic tree is a 6-DoF trunk with four identical 3-link chains, so the branch
// parallelism of every tree algorithm (FK, RNEA, CRBA, the block-arrow factorisation of M) is exactly 4;
// wider cooperative mappings leave lanes idle.  Trunk quantities are replicated in the 4 lanes and the
// per-leg contributions are combined with two __shfl_xor steps (bit-identical in every lane).
//
// Frame convention: world-aligned axes, positions relative to the trunk origin ("base frame B0": an inertial
// frame whose origin coincides with the trunk origin at this instant).  Generalised velocity as in RaiSim:
// gv = [v_trunk_origin (world), omega (world), joint rates]  (ENV:999-1000).
//
// The CPU oracle (oracle/bp5_oracle.hpp) computes the same physics with dense textbook algorithms; this file
// uses composite inertias (CRBA), a recursive Newton-Euler pass for h, and a Schur complement onto the trunk.
#pragma once
#include <cuda_runtime.h>
#include "irrl_params.h"

namespace irrl {

#define IRRL_PI_REF 3.1415926f /* ENV:45 (sic) */
#define FULLMASK 0xffffffffu

// ------------------------------------------------------------------ tiny vector algebra
struct f3 { float x, y, z; };
// IRRL_FASTDIV bits: 1 = inv_sym3, 2 = normal impulse of solve_one_contact, 4 = quaternion increment use __fdividef (MUFU.RCP, 2 ulp)
// instead of IEEE division with its slow-path subroutine; default all three (-3.6 % kernel time, same agreement with the oracle)
#ifndef IRRL_FASTDIV
#define IRRL_FASTDIV 7
#endif
// sine / cosine of the substep loop.  IRRL_SINCOS: 0 = libdevice sincosf, 1 = MUFU (__sincosf, |err| < 4e-7), 2 = sincos_cw below
// (default).  Measured step kernel at 4096 robots with exact divisions: 73.2 / 68.8 / 70.5 us; the MUFU variant flips 5x more
// contact decisions against the fp64 oracle (tests/test_gpu_config_variants.py) and is rejected, sincos_cw agrees like libdevice.
#ifndef IRRL_SINCOS
#define IRRL_SINCOS 2
#endif
// Cody-Waite reduction to [-pi/4, pi/4] + degree-7/8 minimax polynomials: ~1 ulp for |x| < 1e3 (joint angles and half rotation
// increments), branch-free and without the Payne-Hanek slow path that libdevice's sincosf drags into the instruction stream
__device__ __forceinline__ void sincos_cw(float x, float* sp, float* cp) {
    const float k = rintf(x * 0.636619772367581343f);
    float r = fmaf(k, -1.57079601287841796875f, x);          // pi/2 split in three parts (Cody-Waite)
    r = fmaf(k, -3.1391647326017846353352069854736328125e-7f, r);
    r = fmaf(k, -5.390302529957764765544681040410068817436695098876953125e-15f, r);
    const float r2 = r * r;
    float s = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f); s = fmaf(s, r2, -1.6666654611e-1f); s = fmaf(s * r2, r, r);
    float c = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f); c = fmaf(c, r2, 4.166664568298827e-2f); c = fmaf(c * r2, r2, fmaf(r2, -0.5f, 1.0f));
    const int q = (int)k;
    const float s1 = (q & 1) ? c : s, c1 = (q & 1) ? s : c;
    *sp = (q & 2) ? -s1 : s1; *cp = ((q + 1) & 2) ? -c1 : c1;
}
#if IRRL_SINCOS == 0
#define SINCOS(x, s, c) sincosf(x, s, c)
#elif IRRL_SINCOS == 1
#define SINCOS(x, s, c) __sincosf(x, s, c)
#else
#define SINCOS(x, s, c) sincos_cw(x, s, c)
#endif
__device__ __forceinline__ f3 mk(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(f3 a, f3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ f3 axpy(float s, f3 a, f3 b) { return mk(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__device__ __forceinline__ float comp(f3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// MUFU.RSQ alone: rsqrtf() wraps it in a denormal-range fix-up (4 more instructions per call) that the substep loop never needs
// (arguments are squared lengths / pivots of O(1e-6 .. 1e2); every call site refines with one Newton step or guards tiny arguments)
__device__ __forceinline__ float rsq_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// symmetric 3x3
struct S3 { float xx, xy, xz, yy, yz, zz; };
__device__ __forceinline__ f3 mul(const S3& s, f3 v) {
    return mk(fmaf(s.xx, v.x, fmaf(s.xy, v.y, s.xz * v.z)), fmaf(s.xy, v.x, fmaf(s.yy, v.y, s.yz * v.z)),
              fmaf(s.xz, v.x, fmaf(s.yz, v.y, s.zz * v.z)));
}
__device__ __forceinline__ S3 operator+(const S3& a, const S3& b) {
    S3 r; r.xx = a.xx + b.xx; r.xy = a.xy + b.xy; r.xz = a.xz + b.xz; r.yy = a.yy + b.yy; r.yz = a.yz + b.yz; r.zz = a.zz + b.zz; return r;
}
// s += k * e e^T
__device__ __forceinline__ void add_outer(S3& s, float k, f3 e) {
    float kx = k * e.x, ky = k * e.y, kz = k * e.z;
    s.xx = fmaf(kx, e.x, s.xx); s.xy = fmaf(kx, e.y, s.xy); s.xz = fmaf(kx, e.z, s.xz);
    s.yy = fmaf(ky, e.y, s.yy); s.yz = fmaf(ky, e.z, s.yz); s.zz = fmaf(kz, e.z, s.zz);
}
// s += k * (a b^T + b a^T)
__device__ __forceinline__ void add_sym_outer(S3& s, float k, f3 a, f3 b) {
    s.xx = fmaf(2.f * k * a.x, b.x, s.xx); s.yy = fmaf(2.f * k * a.y, b.y, s.yy); s.zz = fmaf(2.f * k * a.z, b.z, s.zz);
    s.xy += k * (a.x * b.y + a.y * b.x); s.xz += k * (a.x * b.z + a.z * b.x); s.yz += k * (a.y * b.z + a.z * b.y);
}
// parallel-axis term m (|c|^2 1 - c c^T)
__device__ __forceinline__ void add_point_mass(S3& s, float m, f3 c) {
    float c2 = dot(c, c);
    s.xx += m * (c2 - c.x * c.x); s.yy += m * (c2 - c.y * c.y); s.zz += m * (c2 - c.z * c.z);
    s.xy -= m * c.x * c.y; s.xz -= m * c.x * c.z; s.yz -= m * c.y * c.z;
}
// inverse of a symmetric PD 3x3
__device__ __forceinline__ S3 inv_sym3(const S3& a) {
    float c00 = a.yy * a.zz - a.yz * a.yz, c01 = a.xz * a.yz - a.xy * a.zz, c02 = a.xy * a.yz - a.xz * a.yy;
    float det = a.xx * c00 + a.xy * c01 + a.xz * c02;
    float id = (IRRL_FASTDIV & 1) ? __fdividef(1.0f, det) : 1.0f / det;           // MUFU.RCP (1 ulp): no slow-path subroutine in the substep loop
    S3 r; r.xx = c00 * id; r.xy = c01 * id; r.xz = c02 * id;
    r.yy = (a.xx * a.zz - a.xz * a.xz) * id; r.yz = (a.xy * a.xz - a.xx * a.yz) * id; r.zz = (a.xx * a.yy - a.xy * a.xy) * id;
    return r;
}

// ------------------------------------------------------------------ packed pairs (FFMA2 / FMUL2 / FADD2 of sm_100a)
// The six trunk dimensions of every trunk-space quantity (B, Y, Q, S, y, ...) are carried as three (even row, odd row) pairs in
// 64-bit register pairs: one FFMA2 does the work of two FFMAs, and a 32-bit operand is broadcast to both halves for free
// (`R.F32` operand form).  Measured on B200 (scripts/microbench/ffma2_latency.cu): dependent FFMA2 4.45 cycles against 4.12 for FFMA,
// issue once per 2.2 cycles per scheduler -- in this latency-bound kernel a packed instruction costs what a scalar one does.
// Each half computes exactly the IEEE operation of the scalar code, in the same order wherever the pair runs over OUTPUT rows.
typedef float2 p2;
__device__ __forceinline__ p2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ p2 fma2s(p2 a, float s, p2 c) { return __ffma2_rn(a, make_float2(s, s), c); }       // a * s + c
__device__ __forceinline__ p2 mul2(p2 a, p2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ p2 mul2s(p2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ p2 add2(p2 a, p2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ p2 sub2(p2 a, p2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float hsum(p2 a) { return a.x + a.y; }
__device__ __forceinline__ float half_of(p2 a, int odd) { return odd ? a.y : a.x; }

// two 3-vectors / two symmetric 3x3 side by side (links 2 and 3 of a leg: same formulas, different data): .x half = first, .y half = second
struct f3p { p2 x, y, z; };
struct S3p { p2 xx, xy, xz, yy, yz, zz; };
__device__ __forceinline__ f3p mkp(f3 a, f3 b) { f3p r; r.x = mk2(a.x, b.x); r.y = mk2(a.y, b.y); r.z = mk2(a.z, b.z); return r; }
__device__ __forceinline__ f3 lo(const f3p& v) { return mk(v.x.x, v.y.x, v.z.x); }
__device__ __forceinline__ f3 hi(const f3p& v) { return mk(v.x.y, v.y.y, v.z.y); }
__device__ __forceinline__ S3 lo(const S3p& s) { return S3{s.xx.x, s.xy.x, s.xz.x, s.yy.x, s.yz.x, s.zz.x}; }
__device__ __forceinline__ S3 hi(const S3p& s) { return S3{s.xx.y, s.xy.y, s.xz.y, s.yy.y, s.yz.y, s.zz.y}; }
__device__ __forceinline__ p2 neg2(p2 a) { return mk2(-a.x, -a.y); }
__device__ __forceinline__ f3p operator+(const f3p& a, const f3p& b) { f3p r; r.x = add2(a.x, b.x); r.y = add2(a.y, b.y); r.z = add2(a.z, b.z); return r; }
// same operation order per half as cross(): t = a.z b.y ; x = a.y b.z - t ; ...
__device__ __forceinline__ f3p crossp(const f3p& a, const f3p& b) {
    f3p r;
    r.x = fma2(a.y, b.z, neg2(mul2(a.z, b.y))); r.y = fma2(a.z, b.x, neg2(mul2(a.x, b.z))); r.z = fma2(a.x, b.y, neg2(mul2(a.y, b.x)));
    return r;
}
__device__ __forceinline__ f3p mulp(const S3p& s, const f3p& v) {      // same order per half as mul(S3, f3)
    f3p r;
    r.x = fma2(s.xx, v.x, fma2(s.xy, v.y, mul2(s.xz, v.z)));
    r.y = fma2(s.xy, v.x, fma2(s.yy, v.y, mul2(s.yz, v.z)));
    r.z = fma2(s.xz, v.x, fma2(s.yz, v.y, mul2(s.zz, v.z)));
    return r;
}
// s = k e e^T (first term of a sum: the scalar code's fmaf(.., 0) is the rounded product) and s += k e e^T, same order as add_outer
__device__ __forceinline__ void set_outerp(S3p& s, p2 k, const f3p& e) {
    const p2 kx = mul2(k, e.x), ky = mul2(k, e.y), kz = mul2(k, e.z);
    s.xx = mul2(kx, e.x); s.xy = mul2(kx, e.y); s.xz = mul2(kx, e.z); s.yy = mul2(ky, e.y); s.yz = mul2(ky, e.z); s.zz = mul2(kz, e.z);
}
__device__ __forceinline__ void add_outerp(S3p& s, p2 k, const f3p& e) {
    const p2 kx = mul2(k, e.x), ky = mul2(k, e.y), kz = mul2(k, e.z);
    s.xx = fma2(kx, e.x, s.xx); s.xy = fma2(kx, e.y, s.xy); s.xz = fma2(kx, e.z, s.xz);
    s.yy = fma2(ky, e.y, s.yy); s.yz = fma2(ky, e.z, s.yz); s.zz = fma2(kz, e.z, s.zz);
}
__device__ __forceinline__ void add_outerps(S3p& s, p2 k, f3 e) {      // the same axis for both halves (e1y)
    const p2 kx = mul2s(k, e.x), ky = mul2s(k, e.y), kz = mul2s(k, e.z);
    s.xx = fma2s(kx, e.x, s.xx); s.xy = fma2s(kx, e.y, s.xy); s.xz = fma2s(kx, e.z, s.xz);
    s.yy = fma2s(ky, e.y, s.yy); s.yz = fma2s(ky, e.z, s.yz); s.zz = fma2s(kz, e.z, s.zz);
}

// ------------------------------------------------------------------ quad (4-lane group) collectives
__device__ __forceinline__ float qsum(float v) {
    v += __shfl_xor_sync(FULLMASK, v, 1); v += __shfl_xor_sync(FULLMASK, v, 2); return v;
}
__device__ __forceinline__ float qmax(float v) {
    v = fmaxf(v, __shfl_xor_sync(FULLMASK, v, 1)); v = fmaxf(v, __shfl_xor_sync(FULLMASK, v, 2)); return v;
}
__device__ __forceinline__ float qbcast(float v, int src_leg) { return __shfl_sync(FULLMASK, v, src_leg, 4); }
__device__ __forceinline__ int qbcasti(int v, int src_leg) { return __shfl_sync(FULLMASK, v, src_leg, 4); }
__device__ __forceinline__ f3 qsum3(f3 v) { return mk(qsum(v.x), qsum(v.y), qsum(v.z)); }
__device__ __forceinline__ p2 qsum2(p2 v) {     // same butterfly as qsum on both halves: 4 shuffles + 2 packed adds
    v = add2(v, mk2(__shfl_xor_sync(FULLMASK, v.x, 1), __shfl_xor_sync(FULLMASK, v.y, 1)));
    return add2(v, mk2(__shfl_xor_sync(FULLMASK, v.x, 2), __shfl_xor_sync(FULLMASK, v.y, 2)));
}
__device__ __forceinline__ p2 qbcast2(p2 v, int src_leg) { return mk2(qbcast(v.x, src_leg), qbcast(v.y, src_leg)); }

// ------------------------------------------------------------------ Philox4x32-10 (same specification as the oracle)
__device__ __forceinline__ uint4 philox(uint32_t seed, uint32_t env, uint32_t tick, uint32_t purpose) {
    uint32_t c0 = env, c1 = tick, c2 = purpose, c3 = 0u, k0 = seed, k1 = 0x1BD11BDAu;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float usym(uint32_t x) { return (float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f; }
__device__ __forceinline__ uint32_t pick(uint4 r, int i) { return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w)); }
// four standard normals (Box-Muller) from one block
__device__ __forceinline__ void gauss4(uint32_t seed, uint32_t env, uint32_t tick, uint32_t purpose, float g[4]) {
    uint4 r = philox(seed, env, tick, purpose);
    float u1 = (float)((r.x >> 8) + 1u) * (1.0f / 16777216.0f), u2 = u01(r.y);
    float rad = sqrtf(-2.0f * logf(u1)), s, c; sincosf(6.283185307179586f * u2, &s, &c);
    g[0] = rad * c; g[1] = rad * s;
    u1 = (float)((r.z >> 8) + 1u) * (1.0f / 16777216.0f); u2 = u01(r.w);
    rad = sqrtf(-2.0f * logf(u1)); sincosf(6.283185307179586f * u2, &s, &c);
    g[2] = rad * c; g[3] = rad * s;
}

// ------------------------------------------------------------------ per-robot / per-leg register state
struct Base {
    f3 p; float qw, qx, qy, qz; f3 v, w;
};
struct LegModel {   // per-env (domain-randomised) leg parameters, ENV:435-477
    float m1, m2, m3, knee_z; f3 com1, com2, com3;
    float sx, sy;   // +-1 : front/hind, left/right
};
struct BaseModel { float m0; f3 com0; float mu, rest, thr; };

// rotation matrix columns from quaternion (w,x,y,z)   ENV:986-992
__device__ __forceinline__ void quat_cols(float w, float x, float y, float z, f3& ex, f3& ey, f3& ez) {
    ex = mk(1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y));
    ey = mk(2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x));
    ez = mk(2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y));
}

// Forward kinematics of one leg (positions relative to the trunk origin, world axes)
struct LegKin {
    f3 a1, a2;            // joint axes (a3 == a2: hip and knee axes are parallel)
    f3 j1, j2, j3, toe;   // joint origins, toe frame origin
    f3 r1, r2, r3;        // COM offsets from the own joint origin
    f3 e1y, e1z, e2x, e2z, e3x, e3z;   // link frame axes needed for the inertia tensors
};
__device__ __forceinline__ void leg_fk(const EnvParams& P, const LegModel& lm, f3 bx, f3 by, f3 bz, f3 q, LegKin& k) {
    float s1, c1, s2, c2, s3, c3;
    SINCOS(q.x, &s1, &c1); SINCOS(q.y, &s2, &c2); SINCOS(q.z, &s3, &c3);
    // R1 = Rb Rx(q1)
    f3 e1x = bx; k.e1y = axpy(c1, by, s1 * bz); k.e1z = axpy(c1, bz, -s1 * by);
    // R2 = R1 Ry(-q2)   (axis 0 -1 0, URDF:79,105)
    k.e2x = axpy(c2, e1x, s2 * k.e1z); k.e2z = axpy(c2, k.e1z, -s2 * e1x);
    // R3 = R2 Ry(-q3)
    k.e3x = axpy(c3, k.e2x, s3 * k.e2z); k.e3z = axpy(c3, k.e2z, -s3 * k.e2x);
    k.a1 = bx; k.a2 = -k.e1y;
    k.j1 = axpy(P.off1x * lm.sx, bx, (P.off1y * lm.sy) * by);
    k.j2 = axpy(P.off2y * lm.sy, k.e1y, k.j1);
    k.j3 = axpy(lm.knee_z, k.e2z, k.j2);
    k.toe = axpy(P.toe_z, k.e3z, k.j3);
    k.r1 = axpy(lm.com1.x, e1x, axpy(lm.com1.y, k.e1y, lm.com1.z * k.e1z));
    k.r2 = axpy(lm.com2.x, k.e2x, axpy(lm.com2.y, k.e1y, lm.com2.z * k.e2z));
    k.r3 = axpy(lm.com3.x, k.e3x, axpy(lm.com3.y, k.e1y, lm.com3.z * k.e3z));
}
// world-frame link inertia tensors about the link COM; links 2 and 3 side by side (I23: .x halves = thigh, .y halves = shank)
__device__ __forceinline__ void leg_inertias(const EnvParams& P, const LegModel& lm, f3 bx, const LegKin& k, S3& I1, S3p& I23) {
    I1 = S3{0, 0, 0, 0, 0, 0}; add_outer(I1, P.I1[0], bx); add_outer(I1, P.I1[1], k.e1y); add_outer(I1, P.I1[2], k.e1z);
    set_outerp(I23, mk2(P.I2[0], P.I3[0]), mkp(k.e2x, k.e3x)); add_outerps(I23, mk2(P.I2[1], P.I3[1]), k.e1y); add_outerp(I23, mk2(P.I2[2], P.I3[2]), mkp(k.e2z, k.e3z));
    S3 I2 = lo(I23);
    add_sym_outer(I2, -P.I2[3] * lm.sy, k.e1y, k.e2z);            // iyz = -0.000228 * sy  (URDF:92,210)
    I23.xx.x = I2.xx; I23.xy.x = I2.xy; I23.xz.x = I2.xz; I23.yy.x = I2.yy; I23.yz.x = I2.yz; I23.zz.x = I2.zz;
}

// Factorised dynamics of one robot, distributed over the quad.
struct Dyn {
    // leg-local (this lane)
    S3 Dinv;              // inverse of the 3x3 leg block of M
    p2 B[3][3];           // coupling block trunk(6) x leg joints(3), columns (P_k ; L_k): B[p][k] = rows (2p, 2p+1) of column k
    p2 Y[3][3];           // B Dinv, same layout
    // trunk, replicated
    p2 LP[3][6];          // Cholesky factor of the Schur complement S: LP[p][c] = rows (2p, 2p+1) of column c (c <= 2p+1), diagonal stored INVERTED
    p2 hb[3];             // bias force on the trunk rows (pairs of rows)
    f3 hl;                // bias force on this leg's joints
    __device__ __forceinline__ float Bv(int a, int k) const { return half_of(B[a >> 1][k], a & 1); }
    __device__ __forceinline__ float Yv(int a, int k) const { return half_of(Y[a >> 1][k], a & 1); }
    __device__ __forceinline__ float hbv(int a) const { return half_of(hb[a >> 1], a & 1); }
    __device__ __forceinline__ float Lv(int i, int j) const { return half_of(LP[i >> 1][j], i & 1); }       // L[i][j], j <= i
};

__device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// Mass matrix blocks + bias forces for the current state.  Mfull (optional) receives this lane's view of the
// un-factorised blocks for the probe kernels: A (21 packed, replicated), B (18), D (6).
__device__ __forceinline__ void dynamics(const EnvParams& P, const LegModel& lm, const BaseModel& bm, const Base& b,
                                         f3 bx, f3 by, f3 bz, const LegKin& k, f3 qd, Dyn& d,
                                         float* Aout = nullptr, S3* Dout = nullptr, bool reduce_hb = true, int leg = 0) {
    S3 I1; S3p I23; leg_inertias(P, lm, bx, k, I1, I23);
    const S3 I2 = lo(I23), I3 = hi(I23);
    const f3 w0 = b.w;
    // ---- velocities
    f3 w1 = axpy(qd.x, k.a1, w0), w2 = axpy(qd.y, k.a2, w1), w3 = axpy(qd.z, k.a2, w2);
    // ---- bias accelerations (joint accelerations and trunk acceleration zero); w2 x a2 == w1 x a2 (hip and knee axes are parallel)
    f3 al1 = qd.x * cross(w0, k.a1);
    const f3 w1xa2 = cross(w1, k.a2);
    f3 al2 = axpy(qd.y, w1xa2, al1);
    f3 al3 = axpy(qd.z, w1xa2, al2);
    f3 d12 = k.j2 - k.j1, d23 = k.j3 - k.j2;
    f3 aj1 = cross(w0, cross(w0, k.j1));
    f3 aj2 = aj1 + cross(al1, d12) + cross(w1, cross(w1, d12));
    f3 aj3 = aj2 + cross(al2, d23) + cross(w2, cross(w2, d23));
    f3 g = mk(0.f, 0.f, P.gravity);
    f3 F1 = lm.m1 * (aj1 + cross(al1, k.r1) + cross(w1, cross(w1, k.r1)) + g);
    f3 N1 = mul(I1, al1) + cross(w1, mul(I1, w1));
    // links 2 and 3 side by side (packed): the same formulas as for link 1, two links per instruction
    const f3p W23 = mkp(w2, w3), AL23 = mkp(al2, al3), R23 = mkp(k.r2, k.r3);
    f3p F23 = (mkp(aj2, aj3) + crossp(AL23, R23)) + crossp(W23, crossp(W23, R23));
    { const p2 m23 = mk2(lm.m2, lm.m3); F23.z = add2(F23.z, mk2(P.gravity, P.gravity)); F23.x = mul2(F23.x, m23); F23.y = mul2(F23.y, m23); F23.z = mul2(F23.z, m23); }
    const f3p N23 = mulp(I23, AL23) + crossp(W23, mulp(I23, W23));
    const f3p RXF = crossp(R23, F23);
    const f3 F2 = lo(F23), F3 = hi(F23), N2 = lo(N23), N3 = hi(N23);
    // ---- backward pass: moments about the joint origins
    f3 n3 = N3 + hi(RXF), f3_ = F3;
    f3 n2 = N2 + lo(RXF) + n3 + cross(d23, f3_), f2_ = F2 + f3_;
    f3 n1 = N1 + cross(k.r1, F1) + n2 + cross(d12, f2_), f1_ = F1 + f2_;
    d.hl = mk(dot(k.a1, n1), dot(k.a2, n2), dot(k.a2, n3));
    f3 fb = f1_, nb = n1 + cross(k.j1, f1_);
    // trunk body itself (identical in the 4 lanes; added after the quad reduction)
    S3 I0 = S3{0, 0, 0, 0, 0, 0}; add_outer(I0, P.I0[0], bx); add_outer(I0, P.I0[1], by); add_outer(I0, P.I0[2], bz);
    f3 r0 = axpy(bm.com0.x, bx, axpy(bm.com0.y, by, bm.com0.z * bz));
    f3 F0 = bm.m0 * (cross(w0, cross(w0, r0)) + g);
    f3 N0 = cross(w0, mul(I0, w0)) + cross(r0, F0);
    // hb_part: this lane's share of the trunk bias wrench (lane 0 also carries the trunk body); callers reduce it together
    // with whatever else they sum over the quad (integrate_substep: one 6-value reduction instead of two)
    if (reduce_hb) { fb = qsum3(fb) + F0; nb = qsum3(nb) + N0; }
    else if (leg == 0) { fb = fb + F0; nb = nb + N0; }
    d.hb[0] = mk2(fb.x, fb.y); d.hb[1] = mk2(fb.z, nb.x); d.hb[2] = mk2(nb.y, nb.z);

    // ---- composite inertias about the trunk origin: (mass, first moment, second moment)
    f3 c3 = k.j3 + k.r3, c2 = k.j2 + k.r2, c1 = k.j1 + k.r1;
    float mC = lm.m3; f3 hC = lm.m3 * c3; S3 IC = I3; add_point_mass(IC, lm.m3, c3);
    // joint 3 (knee): motion subspace (lin = j3 x a, ang = a)
    f3 s3l = cross(k.j3, k.a2);
    f3 P3 = axpy(mC, s3l, cross(k.a2, hC)), L3 = cross(hC, s3l) + mul(IC, k.a2);
    mC += lm.m2; hC = axpy(lm.m2, c2, hC); IC = IC + I2; add_point_mass(IC, lm.m2, c2);
    f3 s2l = cross(k.j2, k.a2);
    f3 P2 = axpy(mC, s2l, cross(k.a2, hC)), L2 = cross(hC, s2l) + mul(IC, k.a2);
    mC += lm.m1; hC = axpy(lm.m1, c1, hC); IC = IC + I1; add_point_mass(IC, lm.m1, c1);
    f3 s1l = cross(k.j1, k.a1);
    f3 P1 = axpy(mC, s1l, cross(k.a1, hC)), L1 = cross(hC, s1l) + mul(IC, k.a1);
    d.B[0][0] = mk2(P1.x, P1.y); d.B[1][0] = mk2(P1.z, L1.x); d.B[2][0] = mk2(L1.y, L1.z);
    d.B[0][1] = mk2(P2.x, P2.y); d.B[1][1] = mk2(P2.z, L2.x); d.B[2][1] = mk2(L2.y, L2.z);
    d.B[0][2] = mk2(P3.x, P3.y); d.B[1][2] = mk2(P3.z, L3.x); d.B[2][2] = mk2(L3.y, L3.z);
    S3 D;   // leg block of M (+ rotor inertias on the diagonal, URDF rotor_inertia)
    D.xx = dot(s1l, P1) + dot(k.a1, L1) + P.rotor[0];
    D.xy = dot(s1l, P2) + dot(k.a1, L2);
    D.xz = dot(s1l, P3) + dot(k.a1, L3);
    D.yy = dot(s2l, P2) + dot(k.a2, L2) + P.rotor[1];
    D.yz = dot(s2l, P3) + dot(k.a2, L3);
    D.zz = dot(s3l, P3) + dot(k.a2, L3) + P.rotor[2];
    if (Dout) *Dout = D;
    d.Dinv = inv_sym3(D);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        d.Y[p][0] = fma2s(d.B[p][2], d.Dinv.xz, fma2s(d.B[p][1], d.Dinv.xy, mul2s(d.B[p][0], d.Dinv.xx)));
        d.Y[p][1] = fma2s(d.B[p][2], d.Dinv.yz, fma2s(d.B[p][1], d.Dinv.yy, mul2s(d.B[p][0], d.Dinv.xy)));
        d.Y[p][2] = fma2s(d.B[p][2], d.Dinv.zz, fma2s(d.B[p][1], d.Dinv.yz, mul2s(d.B[p][0], d.Dinv.xz)));
    }
    // ---- trunk block A (whole-robot composite) and Schur complement S = A - sum_l Y_l B_l^T, reduced over the quad.
    // S2[p][c] = rows (2p, 2p+1) of column c, c <= 2p+1 (the upper half of the diagonal pair c = 2p+1 is carried along unused).
    p2 S2[3][6];
    {
        // this leg's composite contribution to A, in (lin, ang) ordering: [[m 1, -[h]x],[[h]x, Io]]
        S2[0][0] = mk2(mC, 0.f);     S2[0][1] = mk2(0.f, mC);
        S2[1][0] = mk2(0.f, 0.f);    S2[1][1] = mk2(0.f, -hC.z);  S2[1][2] = mk2(mC, hC.y);     S2[1][3] = mk2(0.f, IC.xx);
        S2[2][0] = mk2(hC.z, -hC.y); S2[2][1] = mk2(0.f, hC.x);   S2[2][2] = mk2(-hC.x, 0.f);   S2[2][3] = mk2(IC.xy, IC.xz);
        S2[2][4] = mk2(IC.yy, IC.yz); S2[2][5] = mk2(0.f, IC.zz);
    }
#define SV(i, j) half_of(S2[(i) >> 1][j], (i) & 1)
    if (Aout) {   // un-reduced A for the probes
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) if (j <= i) Aout[tri(i, j)] = SV(i, j);
    }
    // (all loops over the triangle are written with constant bounds and a predicate: nvcc leaves `for (c = 0; c <= a; ++c)` nests partly
    //  rolled, and one run-time index into S / L is enough to push the whole Dyn structure into local memory)
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int c = 0; c < 6; ++c)
            if (c <= 2 * p + 1) S2[p][c] = sub2(S2[p][c], fma2s(d.Y[p][2], d.Bv(c, 2), fma2s(d.Y[p][1], d.Bv(c, 1), mul2s(d.Y[p][0], d.Bv(c, 0)))));
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            if (c < 2 * p + 1) S2[p][c] = qsum2(S2[p][c]);
            else if (c == 2 * p + 1) S2[p][c].y = qsum(S2[p][c].y);
        }
    {   // trunk body
        f3 h0 = bm.m0 * r0; S3 Io = I0; add_point_mass(Io, bm.m0, r0);
        p2 A0[3][6];
        A0[0][0] = mk2(bm.m0, 0.f);   A0[0][1] = mk2(0.f, bm.m0);
        A0[1][0] = mk2(0.f, 0.f);     A0[1][1] = mk2(0.f, -h0.z);  A0[1][2] = mk2(bm.m0, h0.y);  A0[1][3] = mk2(0.f, Io.xx);
        A0[2][0] = mk2(h0.z, -h0.y);  A0[2][1] = mk2(0.f, h0.x);   A0[2][2] = mk2(-h0.x, 0.f);   A0[2][3] = mk2(Io.xy, Io.xz);
        A0[2][4] = mk2(Io.yy, Io.yz); A0[2][5] = mk2(0.f, Io.zz);
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < 6; ++c) if (c <= 2 * p + 1) S2[p][c] = add2(S2[p][c], A0[p][c]);
        if (Aout) {
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) if (j <= i) Aout[tri(i, j)] = qsum(Aout[tri(i, j)]) + half_of(A0[i >> 1][j], i & 1);
        }
    }
    // ---- Cholesky of S (6x6), diagonal stored as its reciprocal.  Column by column; the rows below the pivot are updated as row
    // pairs (packed), every entry with the same products in the same order as the scalar recurrence.
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        float dj = SV(j, j);
#pragma unroll
        for (int c = 0; c < 6; ++c) if (c < j) dj = fmaf(-d.Lv(j, c), d.Lv(j, c), dj);
        float inv = rsq_approx(dj);
        inv = inv * (1.5f - 0.5f * dj * inv * inv);     // one Newton step: full fp32 accuracy
        if (j & 1) d.LP[j >> 1][j].y = inv; else d.LP[j >> 1][j].x = inv;
        if (!(j & 1)) {                                  // the odd row of the pivot's own pair
            float s = SV(j + 1, j);
#pragma unroll
            for (int c = 0; c < 6; ++c) if (c < j) s = fmaf(-d.Lv(j + 1, c), d.Lv(j, c), s);
            d.LP[j >> 1][j].y = s * inv;
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (2 * p > j) {
                p2 s2 = S2[p][j];
#pragma unroll
                for (int c = 0; c < 6; ++c) if (c < j) s2 = fma2s(d.LP[p][c], -d.Lv(j, c), s2);
                d.LP[p][j] = mul2s(s2, inv);
            }
        }
    }
#undef SV
}

// y = L^-1 z  (forward substitution, in place)
__device__ __forceinline__ void fwd6(const Dyn& d, float* z) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float s = z[i];
#pragma unroll
        for (int c = 0; c < 6; ++c) if (c < i) s = fmaf(-d.Lv(i, c), z[c], s);
        z[i] = s * d.Lv(i, i);
    }
}
// the same forward substitution on row pairs, column by column: every row sees the same products in the same order as in fwd6
// (bit-identical), 6 packed + 3 scalar multiply-adds instead of 15
__device__ __forceinline__ void fwd6p(const Dyn& d, p2* z) {
    z[0].x *= d.Lv(0, 0);
    z[0].y = fmaf(-d.Lv(1, 0), z[0].x, z[0].y);
    z[1] = fma2s(d.LP[1][0], -z[0].x, z[1]); z[2] = fma2s(d.LP[2][0], -z[0].x, z[2]);
    z[0].y *= d.Lv(1, 1);
    z[1] = fma2s(d.LP[1][1], -z[0].y, z[1]); z[2] = fma2s(d.LP[2][1], -z[0].y, z[2]);
    z[1].x *= d.Lv(2, 2);
    z[1].y = fmaf(-d.Lv(3, 2), z[1].x, z[1].y);
    z[2] = fma2s(d.LP[2][2], -z[1].x, z[2]);
    z[1].y *= d.Lv(3, 3);
    z[2] = fma2s(d.LP[2][3], -z[1].y, z[2]);
    z[2].x *= d.Lv(4, 4);
    z[2].y = fmaf(-d.Lv(5, 4), z[2].x, z[2].y);
    z[2].y *= d.Lv(5, 5);
}
// x = L^-T y  (back substitution, in place)
__device__ __forceinline__ void bwd6(const Dyn& d, float* y) {
#pragma unroll
    for (int ii = 0; ii < 6; ++ii) {
        const int i = 5 - ii;
        float s = y[i];
#pragma unroll
        for (int c = 0; c < 6; ++c) if (c > i) s = fmaf(-d.Lv(c, i), y[c], s);
        y[i] = s * d.Lv(i, i);
    }
}

// Solve M x = r for a general right-hand side (rb replicated, rl per lane).  Outputs xb (replicated), xl.
__device__ __forceinline__ void solve_full(const Dyn& d, const float* rb, f3 rl, float* xb, f3& xl) {
    f3 t = mul(d.Dinv, rl);
    float z[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) z[a] = qsum(d.Bv(a, 0) * t.x + d.Bv(a, 1) * t.y + d.Bv(a, 2) * t.z);
#pragma unroll
    for (int a = 0; a < 6; ++a) xb[a] = rb[a] - z[a];
    fwd6(d, xb); bwd6(d, xb);
    xl = mk(t.x - (d.Yv(0, 0) * xb[0] + d.Yv(1, 0) * xb[1] + d.Yv(2, 0) * xb[2] + d.Yv(3, 0) * xb[3] + d.Yv(4, 0) * xb[4] + d.Yv(5, 0) * xb[5]),
            t.y - (d.Yv(0, 1) * xb[0] + d.Yv(1, 1) * xb[1] + d.Yv(2, 1) * xb[2] + d.Yv(3, 1) * xb[3] + d.Yv(4, 1) * xb[4] + d.Yv(5, 1) * xb[5]),
            t.z - (d.Yv(0, 2) * xb[0] + d.Yv(1, 2) * xb[1] + d.Yv(2, 2) * xb[2] + d.Yv(3, 2) * xb[3] + d.Yv(4, 2) * xb[4] + d.Yv(5, 2) * xb[5]));
}

// ------------------------------------------------------------------ hard-contact single-contact solve
// v: contact velocity with the current impulse lo applied, G: 3x3 Delassus block (symmetric), n = +z.
// Same rule as the oracle's solve_one_contact (separation / stick / slide with a fixed point on the
// sliding direction).
__device__ __forceinline__ float rsqrt_nr(float x) { x = fmaxf(x, 1e-30f); float r = rsq_approx(x); return r * (1.5f - 0.5f * x * r * r); }   // ~1 ulp
// The sliding direction takes ONE fixed-point step from the stick direction (slide_iters = 1 of the oracle).  Branches, not selects:
// measured 8 us faster at 4096 robots than a select-only version (separating / sticking contacts skip the sliding arithmetic).
__device__ __forceinline__ f3 solve_one_contact(f3 v, const S3& G, const S3& Ginv, f3 lo, float vtn, float mu) {
    f3 e = mk(v.x, v.y, v.z - vtn);
    f3 ls = lo - mul(Ginv, e);
    if (!(ls.z > 0.f)) return mk(0.f, 0.f, 0.f);
    float lt2 = ls.x * ls.x + ls.y * ls.y;
    float mz = mu * ls.z;
    if (lt2 <= mz * mz) return ls;
    f3 b = v - mul(G, lo);
    float il = rsqrt_nr(lt2);
    float dx = ls.x * il, dy = ls.y * il, lnz = 0.f;
    const float rhs = vtn - b.z;
    float den = G.zz + mu * (G.xz * dx + G.yz * dy);
    if (den > 1e-12f) {
        lnz = fmaxf((IRRL_FASTDIV & 2) ? __fdividef(rhs, den) : rhs / den, 0.f);
        float l0 = mu * lnz * dx, l1 = mu * lnz * dy;
        float vx = b.x + G.xx * l0 + G.xy * l1 + G.xz * lnz;
        float vy = b.y + G.xy * l0 + G.yy * l1 + G.yz * lnz;
        float vn2 = vx * vx + vy * vy;
        if (vn2 > 1e-18f) { float iv = rsqrt_nr(vn2); dx = -vx * iv; dy = -vy * iv; }
    }
    den = G.zz + mu * (G.xz * dx + G.yz * dy);
    if (den > 1e-12f) lnz = fmaxf((IRRL_FASTDIV & 2) ? __fdividef(rhs, den) : rhs / den, 0.f);
    return mk(mu * lnz * dx, mu * lnz * dy, lnz);
}

// heightfield lookup (same triangulation as the oracle's Terrain::sample): height and unit normal under (x, y)
__device__ __forceinline__ void terrain_sample(const EnvParams& P, float x, float y, float& height, f3& n) {
    float fx = (x - P.terrain_cx) / P.terrain_dx + 0.5f * (float)(P.terrain_nx - 1), fy = (y - P.terrain_cy) / P.terrain_dy + 0.5f * (float)(P.terrain_ny - 1);
    fx = fminf(fmaxf(fx, 0.f), (float)(P.terrain_nx - 1)); fy = fminf(fmaxf(fy, 0.f), (float)(P.terrain_ny - 1));
    int i = min((int)fx, P.terrain_nx - 2), j = min((int)fy, P.terrain_ny - 2);
    float u = fx - (float)i, v = fy - (float)j;
    const float* h = P.terrain + (size_t)i * P.terrain_ny + j;
    float h00 = __ldg(h), h01 = __ldg(h + 1), h10 = __ldg(h + P.terrain_ny), h11 = __ldg(h + P.terrain_ny + 1);
    float a, b;
    if (u >= v) { a = (h10 - h00) / P.terrain_dx; b = (h11 - h10) / P.terrain_dy; height = h00 + (h10 - h00) * u + (h11 - h10) * v; }
    else        { a = (h11 - h01) / P.terrain_dx; b = (h01 - h00) / P.terrain_dy; height = h00 + (h11 - h01) * u + (h01 - h00) * v; }
    float inv = 1.0f / sqrtf(1.0f + a * a + b * b);
    n = mk(-a * inv, -b * inv, inv);
}
// contact frame rows (t1, t2, n): t1 = x-axis projected on the tangent plane, t2 = n x t1
__device__ __forceinline__ void contact_frame(f3 n, f3& t1, f3& t2) {
    t1 = mk(1.f - n.x * n.x, -n.x * n.y, -n.x * n.z);
    float inv = 1.0f / sqrtf(dot(t1, t1)); t1 = inv * t1;
    t2 = cross(n, t1);
}

// Per-contact data in the reduced trunk space: v_i = c_i + Q_i^T y + T_i lambda_i,  y = sum_j Q_j lambda_j
struct Contact {
    p2 Q[3][3];      // L^-1 E_i^T, Q[p][r] = trunk rows (2p, 2p+1) of contact-frame column r
    S3 T, G, Ginv;   // leg-local Delassus term, full diagonal block and its inverse
    f3 c, lam;
    float vtn;
    int active;
};
// Build Q, G for a contact at point x (relative to the trunk origin).  Jl (3 columns) is the leg part of the
// contact Jacobian (zero for a trunk contact).
// `framed`: rows of the contact Jacobian are expressed in the contact frame (t1, t2, n) instead of world axes (heightfield);
// Jl* must already be rotated by the caller in that case.
__device__ __forceinline__ void contact_setup(const Dyn& d, f3 x, f3 Jl0, f3 Jl1, f3 Jl2, bool on_leg, Contact& ct,
                                              bool framed = false, f3 t1 = f3{1, 0, 0}, f3 t2 = f3{0, 1, 0}, f3 nn = f3{0, 0, 1}) {
    // E = J_b - J_l Y^T : 3x6, rows as pairs of trunk dimensions.  J_b = [1, -[x]x]
    p2 E[3][3];
    E[0][0] = mk2(1.f, 0.f); E[0][1] = mk2(0.f, 0.f);   E[0][2] = mk2(x.z, -x.y);
    E[1][0] = mk2(0.f, 1.f); E[1][1] = mk2(0.f, -x.z);  E[1][2] = mk2(0.f, x.x);
    E[2][0] = mk2(0.f, 0.f); E[2][1] = mk2(1.f, x.y);   E[2][2] = mk2(-x.x, 0.f);
    if (framed) {   // E <- D E with D rows (t1, t2, n)
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const p2 c0 = E[0][p], c1 = E[1][p], c2 = E[2][p];
            E[0][p] = fma2s(c0, t1.x, fma2s(c1, t1.y, mul2s(c2, t1.z)));
            E[1][p] = fma2s(c0, t2.x, fma2s(c1, t2.y, mul2s(c2, t2.z)));
            E[2][p] = fma2s(c0, nn.x, fma2s(c1, nn.y, mul2s(c2, nn.z)));
        }
    }
    if (on_leg) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            E[0][p] = sub2(E[0][p], fma2s(d.Y[p][2], Jl2.x, fma2s(d.Y[p][1], Jl1.x, mul2s(d.Y[p][0], Jl0.x))));
            E[1][p] = sub2(E[1][p], fma2s(d.Y[p][2], Jl2.y, fma2s(d.Y[p][1], Jl1.y, mul2s(d.Y[p][0], Jl0.y))));
            E[2][p] = sub2(E[2][p], fma2s(d.Y[p][2], Jl2.z, fma2s(d.Y[p][1], Jl1.z, mul2s(d.Y[p][0], Jl0.z))));
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        fwd6p(d, E[r]);
#pragma unroll
        for (int p = 0; p < 3; ++p) ct.Q[p][r] = E[r][p];
    }
    // T = J_l Dinv J_l^T
    ct.T = S3{0, 0, 0, 0, 0, 0};
    if (on_leg) {
        // rows of J_l: (Jl0.r, Jl1.r, Jl2.r)
        f3 rx = mk(Jl0.x, Jl1.x, Jl2.x), ry = mk(Jl0.y, Jl1.y, Jl2.y), rz = mk(Jl0.z, Jl1.z, Jl2.z);
        f3 dx = mul(d.Dinv, rx), dy = mul(d.Dinv, ry), dz = mul(d.Dinv, rz);
        ct.T.xx = dot(rx, dx); ct.T.xy = dot(rx, dy); ct.T.xz = dot(rx, dz); ct.T.yy = dot(ry, dy); ct.T.yz = dot(ry, dz); ct.T.zz = dot(rz, dz);
    }
    // G = T + Q^T Q: the three row pairs accumulate in packed form, the two halves are added at the end
    S3 G = ct.T;
    G.xx += hsum(fma2(ct.Q[2][0], ct.Q[2][0], fma2(ct.Q[1][0], ct.Q[1][0], mul2(ct.Q[0][0], ct.Q[0][0]))));
    G.xy += hsum(fma2(ct.Q[2][0], ct.Q[2][1], fma2(ct.Q[1][0], ct.Q[1][1], mul2(ct.Q[0][0], ct.Q[0][1]))));
    G.xz += hsum(fma2(ct.Q[2][0], ct.Q[2][2], fma2(ct.Q[1][0], ct.Q[1][2], mul2(ct.Q[0][0], ct.Q[0][2]))));
    G.yy += hsum(fma2(ct.Q[2][1], ct.Q[2][1], fma2(ct.Q[1][1], ct.Q[1][1], mul2(ct.Q[0][1], ct.Q[0][1]))));
    G.yz += hsum(fma2(ct.Q[2][1], ct.Q[2][2], fma2(ct.Q[1][1], ct.Q[1][2], mul2(ct.Q[0][1], ct.Q[0][2]))));
    G.zz += hsum(fma2(ct.Q[2][2], ct.Q[2][2], fma2(ct.Q[1][2], ct.Q[1][2], mul2(ct.Q[0][2], ct.Q[0][2]))));
    ct.G = G; ct.Ginv = inv_sym3(G);
}

// Q^T z for a trunk-space vector z (three row pairs): packed partial sums, halves added at the end
__device__ __forceinline__ f3 qt_mul(const Contact& ct, const p2* z) {
    return mk(hsum(fma2(ct.Q[2][0], z[2], fma2(ct.Q[1][0], z[1], mul2(ct.Q[0][0], z[0])))),
              hsum(fma2(ct.Q[2][1], z[2], fma2(ct.Q[1][1], z[1], mul2(ct.Q[0][1], z[0])))),
              hsum(fma2(ct.Q[2][2], z[2], fma2(ct.Q[1][2], z[1], mul2(ct.Q[0][2], z[0])))));
}
// Q dl (trunk-space increment of an impulse change), three row pairs
__device__ __forceinline__ void q_mul(const Contact& ct, f3 dl, p2* out) {
#pragma unroll
    for (int p = 0; p < 3; ++p) out[p] = fma2s(ct.Q[p][2], dl.z, fma2s(ct.Q[p][1], dl.y, mul2s(ct.Q[p][0], dl.x)));
}
// One Gauss-Seidel visit of the contact slot owned by lane `owner`: every lane evaluates its own slot, only the
// owner's result is committed and its trunk-space increment is broadcast to the quad.
__device__ __forceinline__ void gs_visit(Contact& ct, p2* y, int leg, int owner, bool frozen, float mu,
                                         float& maxd, float& maxl) {
    // only the owner of an active, unconverged contact evaluates the solve: idle lanes (swinging feet) would otherwise
    // drag the whole warp through the sliding branch with meaningless velocities
    const bool commit = (leg == owner) && ct.active && !frozen;
    f3 dl = mk(0.f, 0.f, 0.f);
    if (commit) {
        f3 v = ct.c + mul(ct.T, ct.lam) + qt_mul(ct, y);
        f3 ln = solve_one_contact(v, ct.G, ct.Ginv, ct.lam, ct.vtn, mu);
        dl = ln - ct.lam;
        ct.lam = ln;
        maxd = fmaxf(maxd, fmaxf(fabsf(dl.x), fmaxf(fabsf(dl.y), fabsf(dl.z))));
        maxl = fmaxf(maxl, fmaxf(fabsf(ln.x), fmaxf(fabsf(ln.y), fabsf(ln.z))));
    }
    p2 dy[3]; q_mul(ct, dl, dy);
#pragma unroll
    for (int p = 0; p < 3; ++p) y[p] = add2(y[p], qbcast2(dy[p], owner));
}

struct ContactOut { int foot_active; f3 foot_impulse; int sweeps; };
// radius of the trunk box about the trunk origin: below this height a corner can touch flat ground (launch-uniform, computed once per kernel)
__device__ __forceinline__ float trunk_box_reach(const EnvParams& P) { return sqrtf(P.box_half[0] * P.box_half[0] + P.box_half[1] * P.box_half[1] + P.box_half[2] * P.box_half[2]); }

// PHASE_SYNC (template flag SYNC): block-wide barriers that keep the warps of a CTA in the same stretch of the substep code (33 KB of
// straight-line SASS against a 32 KB instruction cache), so that one instruction fetch serves all of them.  Four candidate sites per
// substep (start, after the dynamics, before the contact phase, before the state update); IRRL_SYNC_MASK selects them.  Measured with
// 128-thread CTAs: round 1's kernel (55 KB loop) gained 7 % from all four from 8192 robots per GPU; with the round-2 loop (33 KB) one
// barrier per substep -- the one before the contact phase, which also carries the box-contact hand-over vote -- is best (mask 4:
// 82.0 / 141.1 / 237.6 us at 8192 / 16384 / 32768 robots against 83.4 / 143.1 / 239.4 with all four); 4096 robots run without barriers.
#ifndef IRRL_SYNC_MASK
#define IRRL_SYNC_MASK 4
#endif
#define PHASE_SYNC(bit) do { if (SYNC && ((IRRL_SYNC_MASK >> (bit)) & 1)) __syncthreads(); } while (0)
// ------------------------------------------------------------------ one world.integrate() (ENV:768)
// tau: this leg's joint torques.  fext: optional external generalised force on the trunk (6).
// TERR = false compiles the heightfield code out (launch-uniform: flat-ground launches run a kernel without it in the hot loop).
// BOX = false compiles the trunk-box contact path out: when a box corner touches in any robot of the warp (of the CTA with SYNC) the
// call returns false BEFORE touching the state, and the caller repeats the substep with BOX = true (the step kernel keeps a second,
// rarely entered copy of its substep loop for that: 7 KB less code in the hot loop, which is instruction-fetch bound).
template <bool SYNC = false, bool TERR = true, bool BOX = true>
__device__ __forceinline__ bool integrate_substep(const EnvParams& P, const LegModel& lm, const BaseModel& bm, int leg,
                                                  Base& b, f3& q, f3& qd, f3 tau, ContactOut& out, float box_reach) {
    const float dt = P.sim_dt;
    PHASE_SYNC(0);
    f3 bx, by, bz; quat_cols(b.qw, b.qx, b.qy, b.qz, bx, by, bz);
    LegKin k; leg_fk(P, lm, bx, by, bz, q, k);
    Dyn d; dynamics(P, lm, bm, b, bx, by, bz, k, qd, d, nullptr, nullptr, false, leg);
    PHASE_SYNC(1);

#ifdef EXP_PAD   /* experiment: EXP_PAD extra (cheap, 4 independent chains) instructions in the hot loop, to read the slope of time against code size */
    {
        float p0 = qd.x, p1 = qd.y, p2 = qd.z, p3 = tau.x;
#pragma unroll
        for (int i = 0; i < EXP_PAD / 4; ++i) { p0 = fmaf(p0, 1.0001f, 0.5f); p1 = fmaf(p1, 0.9999f, 0.25f); p2 = fmaf(p2, 1.0002f, 0.125f); p3 = fmaf(p3, 0.9998f, 0.0625f); }
        if (p0 + p1 + p2 + p3 == 123.456f) b.p.x += 1.0f;     /* never true; keeps the chains alive */
    }
#endif
    // ---- free acceleration in the factorised form:  t = Dinv r_l,  w = L^-1 (r_b - sum B t)
    f3 rl = mk(tau.x - P.joint_damping * qd.x - d.hl.x, tau.y - P.joint_damping * qd.y - d.hl.y, tau.z - P.joint_damping * qd.z - d.hl.z);
    f3 t = mul(d.Dinv, rl);
    p2 wv[3];      // L^-1 (r_b - sum_l B_l t_l), three row pairs
#pragma unroll
    for (int p = 0; p < 3; ++p) { const p2 v = qsum2(fma2s(d.B[p][2], t.z, fma2s(d.B[p][1], t.y, fma2s(d.B[p][0], t.x, d.hb[p])))); wv[p] = mk2(-v.x, -v.y); }
    fwd6p(d, wv);

    // ---- collision detection against the plane z = 0 (ENV:268) or the heightfield (ENV:264)
    Contact cf;   // foot contact of this leg (slot 0 of this lane)
    const bool terr = TERR && P.terrain != nullptr;
    f3 fn = mk(0.f, 0.f, 1.f), ft1 = mk(1.f, 0.f, 0.f), ft2 = mk(0.f, 1.f, 0.f); float fh = 0.f;
    if (terr) { terrain_sample(P, b.p.x + k.toe.x, b.p.y + k.toe.y, fh, fn); contact_frame(fn, ft1, ft2); }
    f3 xf = axpy(-P.toe_r, fn, k.toe);                                  // contact point on the sphere
    cf.active = ((b.p.z + k.toe.z - fh) * fn.z - P.toe_r <= 0.f) ? 1 : 0;
    f3 Jl0 = cross(k.a1, xf - k.j1), Jl1 = cross(k.a2, xf - k.j2), Jl2 = cross(k.a2, xf - k.j3);
    if (terr) {   // rows of the contact Jacobian in the contact frame
        Jl0 = mk(dot(ft1, Jl0), dot(ft2, Jl0), dot(fn, Jl0)); Jl1 = mk(dot(ft1, Jl1), dot(ft2, Jl1), dot(fn, Jl1)); Jl2 = mk(dot(ft1, Jl2), dot(ft2, Jl2), dot(fn, Jl2));
    }
    // trunk box corners: lane l tests corners 2l and 2l+1, the quad then ranks the hits in corner order
    int hit0 = 0, hit1 = 0; f3 xc0 = mk(0.f, 0.f, 0.f), xc1 = xc0;
    if (terr || __any_sync(FULLMASK, b.p.z <= box_reach)) {
        int c0 = 2 * leg, c1 = 2 * leg + 1;
        f3 l0 = mk((c0 & 1) ? P.box_half[0] : -P.box_half[0], (c0 & 2) ? P.box_half[1] : -P.box_half[1], (c0 & 4) ? P.box_half[2] : -P.box_half[2]);
        f3 l1 = mk((c1 & 1) ? P.box_half[0] : -P.box_half[0], (c1 & 2) ? P.box_half[1] : -P.box_half[1], (c1 & 4) ? P.box_half[2] : -P.box_half[2]);
        xc0 = axpy(l0.x, bx, axpy(l0.y, by, l0.z * bz)); xc1 = axpy(l1.x, bx, axpy(l1.y, by, l1.z * bz));
        if (terr) {
            float h0, h1; f3 n0, n1; terrain_sample(P, b.p.x + xc0.x, b.p.y + xc0.y, h0, n0); terrain_sample(P, b.p.x + xc1.x, b.p.y + xc1.y, h1, n1);
            hit0 = ((b.p.z + xc0.z - h0) * n0.z <= 0.f) ? 1 : 0; hit1 = ((b.p.z + xc1.z - h1) * n1.z <= 0.f) ? 1 : 0;
        } else { hit0 = (b.p.z + xc0.z <= 0.f) ? 1 : 0; hit1 = (b.p.z + xc1.z <= 0.f) ? 1 : 0; }
    }
    int hits = hit0 | (hit1 << 1);
    int allhits = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) allhits |= qbcasti(hits, l) << (2 * l);
    const bool any_box = __any_sync(FULLMASK, allhits != 0);
    const bool any_foot_or_box = __any_sync(FULLMASK, cf.active || allhits != 0);

    p2 ytot2[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) ytot2[p] = mul2s(wv[p], dt);
    f3 lam_leg = mk(0.f, 0.f, 0.f);
    out.sweeps = 0;
    cf.lam = mk(0.f, 0.f, 0.f);
    if (!BOX) {   // hand the substep over to the full version (state untouched so far); the decision is uniform over whatever shares barriers
        if (SYNC ? (__syncthreads_or(any_box ? 1 : 0) != 0) : any_box) return false;
    } else PHASE_SYNC(2);
    if (any_foot_or_box) {
        // ---- foot contact setup (every lane builds its own slot)
        contact_setup(d, xf, Jl0, Jl1, Jl2, true, cf, terr, ft1, ft2, fn);
        {
            f3 vb_ = b.v + cross(b.w, xf);
            if (terr) vb_ = mk(dot(ft1, vb_), dot(ft2, vb_), dot(fn, vb_));
            f3 vpre = vb_ + qd.x * Jl0 + qd.y * Jl1 + qd.z * Jl2;                         // J u (pre-step)
            f3 jt = t.x * Jl0 + t.y * Jl1 + t.z * Jl2;                                    // J_l Dinv r_l
            const f3 qw_ = qt_mul(cf, wv);
            cf.c = vpre + dt * (qw_ + jt);
            cf.vtn = (vpre.z < -bm.thr) ? -bm.rest * vpre.z : 0.f;
        }
        // ---- trunk box contact slot (k-th penetrating corner in corner order belongs to lane k, at most 4)
        Contact cb; cb.active = 0; cb.lam = mk(0.f, 0.f, 0.f);
        if (BOX && any_box) {
            int rank = -1, cnt = 0; f3 xb = mk(0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                // corner c lives on lane c>>1 as xc0 (even) or xc1 (odd)
                f3 src = (c & 1) ? xc1 : xc0;
                float px = qbcast(src.x, c >> 1), py = qbcast(src.y, c >> 1), pz = qbcast(src.z, c >> 1);
                if ((allhits >> c) & 1) { if (cnt == leg) { rank = c; xb = mk(px, py, pz); } cnt++; }
            }
            cb.active = (rank >= 0) ? 1 : 0;
            f3 bn = mk(0.f, 0.f, 1.f), bt1 = mk(1.f, 0.f, 0.f), bt2 = mk(0.f, 1.f, 0.f);
            if (terr) { float bh; terrain_sample(P, b.p.x + xb.x, b.p.y + xb.y, bh, bn); contact_frame(bn, bt1, bt2); }
            contact_setup(d, xb, mk(0, 0, 0), mk(0, 0, 0), mk(0, 0, 0), false, cb, terr, bt1, bt2, bn);
            f3 vpre = b.v + cross(b.w, xb);
            if (terr) vpre = mk(dot(bt1, vpre), dot(bt2, vpre), dot(bn, vpre));
            const f3 qw_ = qt_mul(cb, wv);
            cb.c = vpre + dt * qw_;
            cb.vtn = (vpre.z < -bm.thr) ? -bm.rest * vpre.z : 0.f;
        }
        // ---- per-contact iteration in the 6-dimensional trunk space: feet simultaneously (block Jacobi, they couple only
        //      weakly through the trunk), trunk-box corners one after the other (Gauss-Seidel), same schedule as the oracle
        p2 y[3] = {mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f)};     // sum_j Q_j lambda_j (replicated in the quad)
        bool frozen = !(qsum((float)(cf.active + cb.active)) > 0.f);    // robots without contacts never iterate
        int sweeps = 0;
        // One loop over visits: iteration `it` is sweep `it` while it < jacobi_sweeps (all feet at once, block Jacobi); after that four
        // iterations form one sweep, each with one leg enabled (the feet one after the other, Gauss-Seidel: block Jacobi need not
        // converge when 3-4 feet couple strongly, 0.03 % of bounding contact substeps).  The trunk-box corners and the convergence test
        // run at the end of every sweep.  One copy of the foot code, no inner loop.
        const int J = min(P.jacobi_sweeps, P.solver_iters), it_end = J + 4 * (P.solver_iters - J);
        float maxd = 0.f, maxl = 0.f;
#pragma unroll 1
        for (int it = 0; it < it_end; ++it) {
            if (__all_sync(FULLMASK, frozen)) break;
#ifdef EXP_NO_SEQ
            const bool seq = false;
#else
            const bool seq = it >= J;
#endif
            const int vi = (it - J) & 3;
            {
                f3 dl = mk(0.f, 0.f, 0.f);
                if (cf.active && !frozen && (!seq || leg == vi)) {
                    f3 v = cf.c + mul(cf.T, cf.lam) + qt_mul(cf, y);
                    f3 ln = solve_one_contact(v, cf.G, cf.Ginv, cf.lam, cf.vtn, bm.mu);
                    dl = ln - cf.lam; cf.lam = ln;
                    maxd = fmaxf(maxd, fmaxf(fabsf(dl.x), fmaxf(fabsf(dl.y), fabsf(dl.z))));
                    maxl = fmaxf(maxl, fmaxf(fabsf(ln.x), fmaxf(fabsf(ln.y), fabsf(ln.z))));
                }
                p2 dy[3]; q_mul(cf, dl, dy);
#pragma unroll
                for (int p = 0; p < 3; ++p) y[p] = add2(y[p], qsum2(dy[p]));
            }
            if (seq && vi != 3) continue;                                   // sweep not finished yet
            if (BOX && any_box) {
#pragma unroll 1
                for (int o = 0; o < 4; ++o) {
                    if (!__any_sync(FULLMASK, leg == o && cb.active && !frozen)) continue;
                    gs_visit(cb, y, leg, o, frozen, bm.mu, maxd, maxl);
                }
            }
            maxd = qmax(maxd); maxl = qmax(maxl);
            if (!frozen) { sweeps += 1; if (maxd <= P.solver_tol * maxl) frozen = true; }
            maxd = 0.f; maxl = 0.f;
        }
        out.sweeps = sweeps;
#pragma unroll
        for (int p = 0; p < 3; ++p) ytot2[p] = add2(ytot2[p], y[p]);
        lam_leg = cf.lam;
    }
    out.foot_active = cf.active; out.foot_impulse = cf.lam;
    PHASE_SYNC(3);

    // ---- new velocity: u+ = u + M^-1 (dt r + J^T lambda)
    float ytot[6] = {ytot2[0].x, ytot2[0].y, ytot2[1].x, ytot2[1].y, ytot2[2].x, ytot2[2].y};
    bwd6(d, ytot);                                                      // trunk increment
    f3 jl = mk(dot(Jl0, lam_leg), dot(Jl1, lam_leg), dot(Jl2, lam_leg));   // J_l^T lambda
    f3 tl = axpy(dt, t, mul(d.Dinv, jl));
    const p2 yt0 = mk2(ytot[0], ytot[1]), yt1 = mk2(ytot[2], ytot[3]), yt2 = mk2(ytot[4], ytot[5]);
    f3 dq = mk(tl.x - hsum(fma2(d.Y[2][0], yt2, fma2(d.Y[1][0], yt1, mul2(d.Y[0][0], yt0)))),
               tl.y - hsum(fma2(d.Y[2][1], yt2, fma2(d.Y[1][1], yt1, mul2(d.Y[0][1], yt0)))),
               tl.z - hsum(fma2(d.Y[2][2], yt2, fma2(d.Y[1][2], yt1, mul2(d.Y[0][2], yt0)))));
    b.v = b.v + mk(ytot[0], ytot[1], ytot[2]);
    b.w = b.w + mk(ytot[3], ytot[4], ytot[5]);
    qd = qd + dq;
    // ---- semi-implicit Euler on the configuration
    b.p = axpy(dt, b.v, b.p);
    {
        // |w| and 1/|w| from one MUFU.RSQ + a Newton step (~1 ulp): no IEEE sqrt / division slow paths in the loop
        const float w2 = dot(b.w, b.w), iw = rsqrt_nr(w2), wn = w2 * iw, th = wn * dt; float kk, cw;
        if (th > 1e-8f) { float sh; SINCOS(0.5f * th, &sh, &cw); kk = sh * iw; } else { kk = 0.5f * dt; cw = 1.f; }
        float dx = kk * b.w.x, dy = kk * b.w.y, dz = kk * b.w.z;
        float ow = cw * b.qw - dx * b.qx - dy * b.qy - dz * b.qz;
        float ox = cw * b.qx + dx * b.qw + dy * b.qz - dz * b.qy;
        float oy = cw * b.qy - dx * b.qz + dy * b.qw + dz * b.qx;
        float oz = cw * b.qz + dx * b.qy - dy * b.qx + dz * b.qw;
        float nn = rsqrt_nr(ow * ow + ox * ox + oy * oy + oz * oz);
        b.qw = ow * nn; b.qx = ox * nn; b.qy = oy * nn; b.qz = oz * nn;
    }
    q = axpy(dt, qd, q);
    return true;
}

// ------------------------------------------------------------------ gait generator + IK (ENV:1687-1890), one leg per lane
__device__ __forceinline__ float bezier_b(float ph) { return ph * ph * ph + 3.0f * (ph * ph * (1.0f - ph)); }   // ENV:89

// ENV:1687-1751.  theta starts at 0 (the reference keeps a stale value and prints "error" on infeasible targets).
__device__ __forceinline__ f3 leg_ik(const EnvParams& P, float x, float y, float z, bool is_right) {
    float th0 = 0.f, th1 = 0.f, th2 = 0.f;
    float ll = sqrtf(x * x + y * y + z * z);
    if (ll > P.max_len) { float s = (P.max_len - 1e-5f) / ll; x *= s; y *= s; z *= s; }
    float l_hip = P.l_hip, l_thigh = P.l_thigh, l_calf = P.l_calf;
    float den = z * z + y * y;
    float root = sqrtf(y * y * (den - l_hip * l_hip));
    float temp = is_right ? (-z * l_hip - root) / den : (z * l_hip + root) / den;
    if (fabsf(temp) <= 1.f) th0 = asinf(temp);
    float lr = sqrtf(x * x + y * y + z * z - l_hip * l_hip);
    lr = (lr > (l_thigh + l_calf)) ? (l_thigh + l_calf - 1e-4f) : lr;
    temp = (l_thigh * l_thigh + l_calf * l_calf - lr * lr) / 2.f / l_thigh / l_calf + 1e-5f;
    if (fabsf(temp) <= 1.f) th2 = -(IRRL_PI_REF - acosf(temp));
    float temp1 = x / lr;
    float temp2 = (lr * lr + l_thigh * l_thigh - l_calf * l_calf) / 2.f / lr / l_thigh - 1e-5f;
    if (fabsf(temp1) <= 1.f && fabsf(temp2) <= 1.f) th1 = acosf(temp2) - asinf(temp1);
    return mk(th0, -th1, -th2);   // ENV:1879-1881
}

struct GaitCmd { float gait_step, side_step, rot_step, up_height; };
__device__ __forceinline__ GaitCmd gait_cmd(const EnvParams& P, const float* cf) {
    GaitCmd g;
    g.gait_step = cf[0] * P.lam * P.period; if (P.flag_wildcat) g.gait_step = -g.gait_step;    // ENV:1772-1773
    g.side_step = cf[1] * P.lam * P.period;                                                    // ENV:1775
    g.rot_step = cf[2] * P.period * 0.4f;                                                      // ENV:1777
    g.up_height = P.up_height_max;
    if (P.flag_height_variable) {                                                              // ENV:1779-1792
        float ratio = fabsf(cf[0]) / P.Vx_max;
        if (P.Vy_max > 0.f) ratio = fmaxf(ratio, fabsf(cf[1]) / P.Vy_max);
        if (P.omega_max > 0.f) ratio = fmaxf(ratio, fabsf(cf[2] / P.omega_max));
        g.up_height = (ratio > 0.1f) ? P.up_height_max : ratio * P.up_height_max;
    }
    return g;
}
// reference joint angles + toe target of leg `leg` at time tt (body of the loops ENV:1802-1842 / 1844-1885)
// __noinline__: a single copy of this code serves every call site, so equal inputs give bit-equal joint references
// (the reference relies on that: the second command_obs_update of reset() yields jointDotRef_ == 0 exactly, ENV:628, 1886)
static __device__ __noinline__ f3 leg_reference(const EnvParams& P, const GaitCmd& g, int leg, float tt, f3& toe) {
    float real_phase = fmodf(tt + P.phase[leg] * P.period, P.period) / P.period;
    float anti = (leg < 2) ? 1.0f : -1.0f;
    float ax = g.gait_step / 2.0f, ay = g.side_step / 2.0f + anti * g.rot_step / 2.0f;        // p0 of stance = (ax, ay)
    float bxx = -g.gait_step / 2.0f, byy = -g.side_step / 2.0f + -anti * g.rot_step / 2.0f;   // pf of stance
    if (real_phase < P.lam) {
        float bz = bezier_b(real_phase / P.lam);
        toe = mk(ax + bz * (bxx - ax), ay + bz * (byy - ay), -P.stand_height + bz * 0.f);
    } else {
        float r = (real_phase - P.lam) / (1.0f - P.lam);
        float bz = bezier_b(r);
        float width = 1.0f;
        float gz = g.up_height * expf(-(r - width / 2.f) * (r - width / 2.f) / (2.f * (width / 6.f) * (width / 6.f)));   // ENV:96-99
        toe = mk(bxx + bz * (ax - bxx), byy + bz * (ay - byy), -P.stand_height + gz);
    }
    float off = (leg == 0) ? (-P.l_hip + P.lean_front) : (leg == 1) ? (P.l_hip - P.lean_front) : (leg == 2) ? (-P.l_hip + P.lean_hind) : (P.l_hip - P.lean_hind);   // ENV:1795-1798
    return leg_ik(P, toe.x, toe.y + off, toe.z, (leg == 0) || (leg == 2));
}

// shaping functions for the contact reward (ENV:118-156)
__device__ __forceinline__ float smooth_raw(float phase, float slope, float lam) {
    float f = fmodf(phase, 1.0f);
    if (f < lam) return (sinf(f / lam * 2.f * IRRL_PI_REF) * slope) + 0.5f;
    return (-sinf((f - lam) / (1.0f - lam) * 2.f * IRRL_PI_REF) * slope) + 0.5f;
}
__device__ __forceinline__ float smooth_clamp(float t) { return t > 1.f ? 1.f : (t < 0.f ? 0.f : t); }
__device__ __forceinline__ float smooth_clamp2(float t) { return t > 1.f ? 0.f : (t < 0.f ? 1.f : 1.f - t); }
__device__ __forceinline__ float smooth_function(float phase, float slope, float lam) { return smooth_clamp(smooth_raw(phase, slope, lam)); }
__device__ __forceinline__ float smooth_function2(float phase, float slope, float lam) { return smooth_clamp2(smooth_raw(phase, slope, lam)); }

}  // namespace irrl
