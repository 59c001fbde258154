// =====================================================================================
// bp5_oracle.hpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH)
//
// A plain C++ restatement of the reference's vectorised bp5 environment step, used only as
// the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs.  The product (CUDA) path never includes, links or calls anything in oracle/.
//
// Citations are file:line under /root/reference/IRRL/FlexibleRobotRaisimGym/flex_gym/env:
//   ENV = env/BlackPanther_V55/Environment.hpp     VEC = VectorizedEnvironment.hpp
//   URDF = env/BlackPanther_V55/urdf/black_panther.urdf
//
// PARITY STATUS
//   * Everything outside world_->integrate() (action scaling, PD, torque clamp, observation,
//     reward, command sampling, gait generator + IK, termination, reset order, auto-reset)
//     follows ENV / VEC line by line.  The gait generator + IK are PINNED by the reference's
//     shipped trajectory dump Exp_Raw_Data/trot_ref_.csv (tests/test_oracle_gait_kat.py).
//   * world_->integrate() lives in RaiSim, a closed-source third-party library that is absent
//     from /root/reference and from this image (CMakeLists.txt:13 find_package(raisimOgre 0.6.0);
//     readme "Raisim >= 1.0.0").  Its published algorithm (Hwangbo, Lee, Hutter, "Per-contact
//     iteration method for solving contact dynamics", RA-L 2018; RaiSim manual conventions for
//     gc/gv) is restated here: M(q) and h(q,u) are uniquely defined by the URDF, the contact
//     solver is a per-contact Gauss-Seidel over hard contacts with Coulomb friction and Newton
//     restitution, semi-implicit Euler integration.  There is no RaiSim output to diff against:
//     for that boundary the oracle is "PARITY UNPINNED" and is checked only by analytic
//     invariants (tests/test_oracle_dynamics.py: M vs kinetic-energy finite differences, energy
//     and momentum conservation, static force balance).
//
// The formulation here is deliberately the textbook dense one (explicit body Jacobians, dense
// 18x18 Cholesky, dense Delassus matrix) so that it is structurally independent from the CUDA
// kernel's leg-lane composite-inertia / Schur-complement formulation.
// =====================================================================================
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <stdexcept>

namespace bp5o {

// ------------------------------------------------------------------ scalar functions (mth::) and the op-counting scalar
// The env is templated on its scalar type; Counted<double> counts every floating-point operation it performs ("the
// oracle carries an op counter"): add/sub, mul, div, sqrt, transcendental, compare.  bp5o_count_flops() reports them per env-step.
struct OpCount { unsigned long long add = 0, mul = 0, div = 0, sqrt = 0, trans = 0, cmp = 0; };
inline OpCount& op_count() { static thread_local OpCount c; return c; }
struct Counted {
    double v;
    Counted() : v(0) {}
    Counted(double x) : v(x) {}
    explicit operator double() const { return v; }
    explicit operator float() const { return (float)v; }
    explicit operator int() const { return (int)v; }
    friend Counted operator+(const Counted& a, const Counted& b) { op_count().add++; return Counted(a.v + b.v); }
    friend Counted operator-(const Counted& a, const Counted& b) { op_count().add++; return Counted(a.v - b.v); }
    friend Counted operator*(const Counted& a, const Counted& b) { op_count().mul++; return Counted(a.v * b.v); }
    friend Counted operator/(const Counted& a, const Counted& b) { op_count().div++; return Counted(a.v / b.v); }
    Counted operator-() const { return Counted(-v); }
    Counted& operator+=(const Counted& b) { op_count().add++; v += b.v; return *this; }
    Counted& operator-=(const Counted& b) { op_count().add++; v -= b.v; return *this; }
    Counted& operator*=(const Counted& b) { op_count().mul++; v *= b.v; return *this; }
    Counted& operator/=(const Counted& b) { op_count().div++; v /= b.v; return *this; }
    friend bool operator<(const Counted& a, const Counted& b) { op_count().cmp++; return a.v < b.v; }
    friend bool operator>(const Counted& a, const Counted& b) { op_count().cmp++; return a.v > b.v; }
    friend bool operator<=(const Counted& a, const Counted& b) { op_count().cmp++; return a.v <= b.v; }
    friend bool operator>=(const Counted& a, const Counted& b) { op_count().cmp++; return a.v >= b.v; }
    friend bool operator==(const Counted& a, const Counted& b) { op_count().cmp++; return a.v == b.v; }
    friend bool operator!=(const Counted& a, const Counted& b) { op_count().cmp++; return a.v != b.v; }
};
namespace mth {
using std::sqrt; using std::fabs; using std::exp; using std::sin; using std::cos; using std::min; using std::max; using std::fmod; using std::fmax;
using std::fmin; using std::asin; using std::acos; using std::log; using std::atan2;
inline Counted sqrt(const Counted& a) { op_count().sqrt++; return Counted(std::sqrt(a.v)); }
inline Counted fabs(const Counted& a) { return Counted(std::fabs(a.v)); }
inline Counted exp(const Counted& a) { op_count().trans++; return Counted(std::exp(a.v)); }
inline Counted sin(const Counted& a) { op_count().trans++; return Counted(std::sin(a.v)); }
inline Counted cos(const Counted& a) { op_count().trans++; return Counted(std::cos(a.v)); }
inline Counted asin(const Counted& a) { op_count().trans++; return Counted(std::asin(a.v)); }
inline Counted acos(const Counted& a) { op_count().trans++; return Counted(std::acos(a.v)); }
inline Counted log(const Counted& a) { op_count().trans++; return Counted(std::log(a.v)); }
inline Counted atan2(const Counted& a, const Counted& b) { op_count().trans++; return Counted(std::atan2(a.v, b.v)); }
inline Counted fmod(const Counted& a, const Counted& b) { op_count().div++; return Counted(std::fmod(a.v, b.v)); }
inline Counted min(const Counted& a, const Counted& b) { op_count().cmp++; return a.v < b.v ? a : b; }
inline Counted max(const Counted& a, const Counted& b) { op_count().cmp++; return a.v < b.v ? b : a; }
inline Counted fmin(const Counted& a, const Counted& b) { op_count().cmp++; return a.v < b.v ? a : b; }
inline Counted fmax(const Counted& a, const Counted& b) { op_count().cmp++; return a.v < b.v ? b : a; }
}  // namespace mth

// ------------------------------------------------------------------ Philox4x32-10 (Salmon et al. SC'11)
// Counter-based RNG shared *by specification* with the CUDA path: key = (seed, 0x1BD11BDA),
// counter = (global env id, tick, purpose, 0).  Replaces the reference's racy libc rand()/random()
// and Eigen setRandom (ENV:440-476, 557-605, 704, 977-1003, 1027-1075), which are not reproducible
// (SURVEY 9.3 quirk 6); only the distributions are kept.
struct Philox {
    static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = (uint64_t)a * (uint64_t)b; hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
    }
    static inline void gen(uint32_t seed, uint32_t env, uint32_t tick, uint32_t purpose, uint32_t out[4]) {
        uint32_t c0 = env, c1 = tick, c2 = purpose, c3 = 0u;
        uint32_t k0 = seed, k1 = 0x1BD11BDAu;
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};
// uniform in [0,1): 24 bits, exactly representable in fp32
static inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// uniform in [-1,1)  (Eigen setRandom distribution)
static inline float usym(uint32_t x) { return (float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f; }

// RNG purposes (tick-local stream ids); +32 selects the "inside reset()" copy of the same draw.
enum : uint32_t {
    P_OBS_Q0 = 0, P_OBS_Q1 = 1, P_OBS_Q2 = 2, P_OBS_QD0 = 3, P_OBS_QD1 = 4, P_OBS_QD2 = 5,
    P_OBS_POSTURE = 6, P_OBS_OMEGA = 7, P_CMD = 8, P_ACT = 9,
    P_RST_TIME_CMD = 16, P_RST_INIT = 17, P_RST_BASEVEL = 18, P_RST_XY = 19, P_DISTURB = 20,
    P_IN_RESET = 32,
    P_DR_MATERIAL = 64, P_DR_MASS = 65 /* +body (13) */, P_DR_COM = 80 /* +body (13) */, P_DR_CALF = 96,
};

// ------------------------------------------------------------------ small linear algebra
template <typename T> struct V3 {
    T x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(T a, T b, T c) : x(a), y(b), z(c) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename T> V3<T> operator+(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> V3<T> operator-(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> V3<T> operator*(T s, const V3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T> T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> V3<T> cross(const V3<T>& a, const V3<T>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T> struct M3 {
    T m[3][3];
    M3() { for (auto& r : m) for (auto& e : r) e = T(0); }
    static M3 eye() { M3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = T(1); return r; }
    V3<T> col(int j) const { return {m[0][j], m[1][j], m[2][j]}; }
};
template <typename T> M3<T> operator*(const M3<T>& a, const M3<T>& b) {
    M3<T> r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T s = 0; for (int k = 0; k < 3; ++k) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
    return r;
}
template <typename T> V3<T> operator*(const M3<T>& a, const V3<T>& v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
template <typename T> M3<T> transpose(const M3<T>& a) { M3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
template <typename T> V3<T> tmul(const M3<T>& a, const V3<T>& v) {  // a^T v
    return {a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z,
            a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z,
            a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z};
}
// Rodrigues rotation about a unit axis
template <typename T> M3<T> axis_angle(const V3<T>& a, T th) {
    T c = mth::cos(th), s = mth::sin(th), v = T(1) - c;
    M3<T> r;
    r.m[0][0] = c + a.x * a.x * v;       r.m[0][1] = a.x * a.y * v - a.z * s; r.m[0][2] = a.x * a.z * v + a.y * s;
    r.m[1][0] = a.y * a.x * v + a.z * s; r.m[1][1] = c + a.y * a.y * v;       r.m[1][2] = a.y * a.z * v - a.x * s;
    r.m[2][0] = a.z * a.x * v - a.y * s; r.m[2][1] = a.z * a.y * v + a.x * s; r.m[2][2] = c + a.z * a.z * v;
    return r;
}
// quaternion (w,x,y,z) -> rotation matrix (ENV:986-992 raisim::quatToRotMat, standard formula)
template <typename T> M3<T> quat_to_rot(const T q[4]) {
    T w = q[0], x = q[1], y = q[2], z = q[3];
    M3<T> r;
    r.m[0][0] = T(1) - T(2) * (y * y + z * z); r.m[0][1] = T(2) * (x * y - w * z);        r.m[0][2] = T(2) * (x * z + w * y);
    r.m[1][0] = T(2) * (x * y + w * z);        r.m[1][1] = T(1) - T(2) * (x * x + z * z); r.m[1][2] = T(2) * (y * z - w * x);
    r.m[2][0] = T(2) * (x * z - w * y);        r.m[2][1] = T(2) * (y * z + w * x);        r.m[2][2] = T(1) - T(2) * (x * x + y * y);
    return r;
}

// ------------------------------------------------------------------ robot model (URDF restated)
constexpr int NB = 13;   // bodies after fixed-joint merge (ENV:449 loops 13 bodies)
constexpr int NV = 18;   // gvDim  ENV:296
constexpr int NQ = 19;   // gcDim  ENV:295
constexpr int NJ = 12;
constexpr int MAXC = 8;  // 4 toe spheres + up to 4 trunk-box corners

template <typename T> struct Body {
    int parent;       // body index, -1 for the trunk
    V3<T> off;        // joint origin in parent frame
    V3<T> axis;       // joint axis in own/parent frame (revolute)
    T mass;
    V3<T> com;        // in body frame
    M3<T> I;          // about COM, body frame
    T rotor;          // rotor inertia reflected on the joint (added to M diagonal)
};

template <typename T> struct Model {
    Body<T> b[NB];
    V3<T> toe_off;       // toe_*_joint origin in shank frame  URDF:162
    T toe_radius;        // URDF:148
    V3<T> box_half;      // trunk collision box half extents   URDF:26
    T gravity;
    T joint_damping;     // URDF:56 <dynamics damping="0.01"> (switch; RaiSim-side behaviour unpinned)
    Model() {
        // trunk  URDF:9-29 (+ zero-mass dummy link URDF:31-47)
        b[0].parent = -1; b[0].mass = T(3.72); b[0].com = {T(0), T(0), T(-0.003)};
        b[0].I.m[0][0] = T(0.016269); b[0].I.m[1][1] = T(0.050813); b[0].I.m[2][2] = T(0.060989);
        b[0].rotor = 0;
        const int sxs[4] = {+1, +1, -1, -1};  // FR FL HR HL : front/hind   URDF:52,171,289,406
        const int sys[4] = {-1, +1, -1, +1};  //               right/left
        for (int l = 0; l < 4; ++l) {
            T sx = T(sxs[l]), sy = T(sys[l]);
            Body<T>& a = b[1 + 3 * l];  // abduct_*  URDF:50-76
            a.parent = 0; a.off = {T(0.212) * sx, T(0.051) * sy, T(0)}; a.axis = {T(1), T(0), T(0)};
            a.mass = T(0.54); a.com = {T(0.058) * sx, T(0.00485) * sy, T(0)};
            a.I.m[0][0] = T(0.000391); a.I.m[1][1] = T(0.000739); a.I.m[2][2] = T(0.000488);
            a.rotor = T(0.003708);
            Body<T>& t = b[2 + 3 * l];  // thigh_*   URDF:78-102
            t.parent = 1 + 3 * l; t.off = {T(0), T(0.085) * sy, T(0)}; t.axis = {T(0), T(-1), T(0)};
            t.mass = T(0.636); t.com = {T(0), T(-0.019) * sy, T(-0.01865)};
            t.I.m[0][0] = T(0.001724); t.I.m[1][1] = T(0.001907); t.I.m[2][2] = T(0.000468);
            t.I.m[1][2] = t.I.m[2][1] = T(-0.000228) * sy;
            t.rotor = T(0.003708);
            Body<T>& s = b[3 + 3 * l];  // shank_* merged with toe_* through the fixed toe joint URDF:104-165
            s.parent = 2 + 3 * l; s.off = {T(0), T(0), T(-0.201)}; s.axis = {T(0), T(-1), T(0)};
            const T ms = T(0.064), mt = T(0.05), zs = T(-0.0865), zt = T(-0.19);
            s.mass = ms + mt;
            T zc = (ms * zs + mt * zt) / (ms + mt);
            s.com = {T(0), T(0), zc};
            T ds = zs - zc, dt = zt - zc;
            s.I.m[0][0] = T(0.000716) + ms * ds * ds + T(0.000025) + mt * dt * dt;
            s.I.m[1][1] = T(0.000721) + ms * ds * ds + T(0.000025) + mt * dt * dt;
            s.I.m[2][2] = T(0.000012) + T(0.000025);
            s.rotor = T(0.008966);
        }
        toe_off = {T(0), T(0), T(-0.19)};
        toe_radius = T(0.0275);
        box_half = {T(0.15), T(0.1), T(0.05)};
        gravity = T(9.81);
        joint_damping = T(0.01);
    }
    // A robot description other than the shipped one, as the compact mirror-symmetric model of include/irrl_b200.h (irrl_parse_urdf):
    // I0(3) I1(3) I2(xx,yy,zz,|yz|) I3(3) rotor(3) off1x off1y off2y toe_z toe_r box_half(3) joint_damping m0 com0(3) m1 com1(3) m2 com2(3) m3 com3z knee_z
    void set_compact(const double* m) {
        const double *I0 = m, *I1 = m + 3, *I2 = m + 6, *I3 = m + 10, *rot = m + 13, *box = m + 21, *c0 = m + 26, *c1 = m + 30, *c2 = m + 34;
        const double off1x = m[16], off1y = m[17], off2y = m[18], toe_z = m[19], toe_r = m[20], damp = m[24], m0 = m[25], m1 = m[29], m2 = m[33], m3 = m[37], com3z = m[38], knee_z = m[39];
        b[0].mass = T(m0); b[0].com = {T(c0[0]), T(c0[1]), T(c0[2])};
        b[0].I = M3<T>(); b[0].I.m[0][0] = T(I0[0]); b[0].I.m[1][1] = T(I0[1]); b[0].I.m[2][2] = T(I0[2]);
        const int sxs[4] = {+1, +1, -1, -1}, sys[4] = {-1, +1, -1, +1};
        for (int l = 0; l < 4; ++l) {
            T sx = T(sxs[l]), sy = T(sys[l]);
            Body<T>& a = b[1 + 3 * l]; a.off = {T(off1x) * sx, T(off1y) * sy, T(0)}; a.mass = T(m1); a.com = {T(c1[0]) * sx, T(c1[1]) * sy, T(c1[2])};
            a.I = M3<T>(); a.I.m[0][0] = T(I1[0]); a.I.m[1][1] = T(I1[1]); a.I.m[2][2] = T(I1[2]); a.rotor = T(rot[0]);
            Body<T>& t = b[2 + 3 * l]; t.off = {T(0), T(off2y) * sy, T(0)}; t.mass = T(m2); t.com = {T(c2[0]) * sx, T(c2[1]) * sy, T(c2[2])};
            t.I = M3<T>(); t.I.m[0][0] = T(I2[0]); t.I.m[1][1] = T(I2[1]); t.I.m[2][2] = T(I2[2]); t.I.m[1][2] = t.I.m[2][1] = T(-I2[3]) * sy; t.rotor = T(rot[1]);
            Body<T>& s = b[3 + 3 * l]; s.off = {T(0), T(0), T(knee_z)}; s.mass = T(m3); s.com = {T(0), T(0), T(com3z)};
            s.I = M3<T>(); s.I.m[0][0] = T(I3[0]); s.I.m[1][1] = T(I3[1]); s.I.m[2][2] = T(I3[2]); s.rotor = T(rot[2]);
        }
        toe_off = {T(0), T(0), T(toe_z)}; toe_radius = T(toe_r); box_half = {T(box[0]), T(box[1]), T(box[2])}; joint_damping = T(damp);
    }
};

// ------------------------------------------------------------------ configuration (YAML keys of ENV:1594-1659, VEC:146-171)
struct Cfg {
    std::map<std::string, double> kv;
    double get(const std::string& k) const {
        auto it = kv.find(k);
        if (it == kv.end()) throw std::runtime_error("Node cfg[\"" + k + "\"] doesn't exist");  // GYM:41-42 READ_YAML
        return it->second;
    }
    double get_or(const std::string& k, double d) const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
    bool flag(const std::string& k) const { return get(k) != 0.0; }
    // "key=value;key=value" (booleans as 0/1)
    static Cfg parse(const char* s) {
        Cfg c; std::string str(s); size_t pos = 0;
        while (pos < str.size()) {
            size_t e = str.find(';', pos); if (e == std::string::npos) e = str.size();
            std::string item = str.substr(pos, e - pos); pos = e + 1;
            size_t eq = item.find('='); if (eq == std::string::npos) continue;
            c.kv[item.substr(0, eq)] = std::stod(item.substr(eq + 1));
        }
        return c;
    }
};

// ------------------------------------------------------------------ helper shape functions
#define BP5O_PI 3.1415926  /* ENV:45 (sic) */
inline double sampling_reshape(double ratio) { return (ratio < 0.5 && ratio > 0) ? ratio * 4.0 / 3.0 : (2.0 * ratio + 1.0) / 3.0; }   // ENV:71-81
template <typename T> inline T bezier_b(T ph) { return ph * ph * ph + T(3.0) * (ph * ph * (T(1.0) - ph)); }  // ENV:89
template <typename T> inline V3<T> cubicBezier(const V3<T>& p0, const V3<T>& pf, T ph) {                      // ENV:86-91
    T b = bezier_b(ph); return p0 + b * (pf - p0);
}
template <typename T> inline T gauss(T x, T width, T height) {                                               // ENV:96-99
    return height * mth::exp(-(x - width / 2) * (x - width / 2) / (2 * (width / 6) * (width / 6)));
}
template <typename T> inline V3<T> Bezier2(const V3<T>& p0, const V3<T>& pf, T ph, T height) {                // ENV:104-113
    T b = bezier_b(ph);
    return {p0.x + b * (pf.x - p0.x), p0.y + b * (pf.y - p0.y), p0.z + gauss(ph, T(1.0), height)};
}
template <typename T> inline T smooth_raw(T phase, T slope, T lam) {                                         // ENV:118-129
    T f = mth::fmod(phase, T(1.0));
    if (f < lam) return (mth::sin(f / lam * 2 * T(BP5O_PI)) * slope) + T(0.5);
    return (-mth::sin((f - lam) / (T(1.0) - lam) * 2 * T(BP5O_PI)) * slope) + T(0.5);
}
template <typename T> inline T smooth_function(T phase, T slope, T lam) {                                    // ENV:118-136
    T t = smooth_raw(phase, slope, lam); if (t > T(1)) return T(1); if (t < T(0)) return T(0); return t;
}
template <typename T> inline T smooth_function2(T phase, T slope, T lam) {                                   // ENV:138-156
    T t = smooth_raw(phase, slope, lam); if (t > T(1)) return T(0); if (t < T(0)) return T(1); return T(1) - t;
}

// ------------------------------------------------------------------ dense Cholesky
template <typename T, int N> struct Chol {
    T L[N][N];
    bool factor(const T A[N][N]) {
        for (int j = 0; j < N; ++j) {
            T d = A[j][j]; for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
            if (!(d > T(0))) return false;
            L[j][j] = mth::sqrt(d);
            for (int i = j + 1; i < N; ++i) { T s = A[i][j]; for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k]; L[i][j] = s / L[j][j]; }
            for (int i = 0; i < j; ++i) L[i][j] = T(0);
        }
        return true;
    }
    void solve(const T* b, T* x) const {
        T y[N];
        for (int i = 0; i < N; ++i) { T s = b[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
        for (int i = N - 1; i >= 0; --i) { T s = y[i]; for (int k = i + 1; k < N; ++k) s -= L[k][i] * x[k]; x[i] = s / L[i][i]; }
    }
};
template <typename T> inline bool inv3(const T A[3][3], T R[3][3]) {
    T c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2], c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    T det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    if (det == T(0)) return false;
    T id = T(1) / det;
    R[0][0] = c00 * id; R[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id; R[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    R[1][0] = c01 * id; R[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id; R[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    R[2][0] = c02 * id; R[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id; R[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
    return true;
}

// ------------------------------------------------------------------ heightfield terrain (ENV:252-265 RaiSim HeightMap; new spec in DESIGN.md)
// Regular grid of nx x ny samples covering x_size x y_size metres centred on (cx, cy); every cell is split along the
// (i,j)-(i+1,j+1) diagonal into two triangles; outside the grid the border value continues.
struct Terrain {
    const float* h = nullptr; int nx = 0, ny = 0; double cx = 0, cy = 0, dx = 1, dy = 1;
    bool valid() const { return h != nullptr; }
    template <typename T> void sample(T x, T y, T& height, V3<T>& n) const {
        T fx = (x - T(cx)) / T(dx) + T(0.5) * T(nx - 1), fy = (y - T(cy)) / T(dy) + T(0.5) * T(ny - 1);   // centre-relative: small magnitudes in fp32
        if (fx < T(0)) fx = T(0); if (fy < T(0)) fy = T(0);
        if (fx > T(nx - 1)) fx = T(nx - 1); if (fy > T(ny - 1)) fy = T(ny - 1);
        int i = (int)fx, j = (int)fy; if (i > nx - 2) i = nx - 2; if (j > ny - 2) j = ny - 2;
        T u = fx - T(i), v = fy - T(j);
        T h00 = T(h[(size_t)i * ny + j]), h10 = T(h[(size_t)(i + 1) * ny + j]), h01 = T(h[(size_t)i * ny + j + 1]), h11 = T(h[(size_t)(i + 1) * ny + j + 1]);
        T a, b;   // slopes dh/dx, dh/dy of the triangle under (u,v)
        if (u >= v) { a = (h10 - h00) / T(dx); b = (h11 - h10) / T(dy); height = h00 + (h10 - h00) * u + (h11 - h10) * v; }
        else        { a = (h11 - h01) / T(dx); b = (h01 - h00) / T(dy); height = h00 + (h11 - h01) * u + (h01 - h00) * v; }
        T inv = T(1) / mth::sqrt(T(1) + a * a + b * b);
        n = V3<T>(-a * inv, -b * inv, inv);
    }
};

// ------------------------------------------------------------------ one robot ("ENVIRONMENT", ENV:212)
template <typename T> struct Env {
    // ---- configuration (ENV:1594-1659)
    T abad_, period_, lam_, stand_height_, up_height_, up_height_max_, down_height_, gait_step_, side_step_ = 0, rot_step_ = 0;
    T Vx_max, Vx_min = 0, Vy_max, Vy_min, omega_max, omega_min, Lean_middle_front, Lean_middle_hind;
    bool flag_terrain, flag_manual, flag_crucial, flag_filter, flag_StochasticDynamics, flag_HeightVariable,
        flag_TimeBasedContact, flag_ManualTraj, flag_MotorDynamics, flag_ObsFilter, flag_WildCat, flag_ForceDisturbance;
    T terminalRewardCoeff_, EECoeff, BodyPosCoeff, BodyAttiCoeff, JointMimicCoeff, VelKeepCoeff, TorqueCoeff, ContactCoeff;
    T stiffness, abad_ratio, damping, freq, max_time, actionNoise, noise_flag;
    int gaitType;
    T MotorMaxTorque, MotorCriticalSpeed, MotorMaxSpeed;
    T simulation_dt_, control_dt_;
    // hard-coded members (ENV:1929-2083)
    T l_thigh_ = T(0.209), l_calf_ = T(0.2175), l_hip_ = T(0.085), max_len = 0;   // ENV:1949-1953
    T jointNoise = T(0.002), jointVelocityNoise = T(0.8);                          // ENV:1988-1989
    T noise_posture_sigma = T(0.02), noise_omega_sigma = T(0.5);                   // ENV:1997,1999
    T cmd_update_param = T(0.995);                                                 // ENV:2043
    T ObsFilterFreq = 20, ObsFilterAlpha = 1;                                      // ENV:2026-2027
    T mass_distrubance_ratio = T(0.15), com_distrubance = T(0.02), calf_distrubance = T(0.01);  // ENV:2069-2071
    // contact material: default (0.6, 0.2, 0.01) ENV:433
    T mu = T(0.6), restitution = T(0.2), rest_threshold = T(0.01);
    int solver_iters = 30;        // per-contact Gauss-Seidel sweeps cap (new spec)
    T solver_tol = T(1e-5);       // relative impulse change for early exit
    int slide_iters = 1;          // fixed-point iterations for the sliding direction
    int solver_jacobi = 1;        // 1: feet updated simultaneously per sweep (see integrate())
    int jacobi_sweeps = 10;       // sweeps after which the feet are visited one after the other too (block Jacobi need not converge when 3-4 feet couple strongly)
    int slide_exact = 0;          // 1: exact maximal-dissipation sliding solve by bisection on the cone boundary (RaiSim's rule, Hwangbo et al. 2018);
                                  // CPU-only yardstick for the one-step direction update the product uses (tests/test_oracle_contact_solver.py)

    // ---- state
    Model<T> model;
    uint32_t seed = 1, env_id = 0, tick = 0;
    T gc_[NQ], gv_[NV], gc_init_[NQ];
    T pTarget12_[NJ], pTarget12Last_[NJ], actionMean_[NJ];
    T torque[NJ], torque_last[NJ], torque_limit[NJ];
    T gf_[NV];                     // last applied generalized force (probe GetGeneralizedForce ENV:1363)
    T jointPgain[NJ], jointDgain[NJ];
    T obDouble_[35], obDouble_last_[35], obMean_[35], obStd_[35];
    T bodyLinearVel_[3], bodyAngularVel_[3];
    M3<T> bodyFrameMatrix_;
    T phase_[4];
    T t0_ = 0; int frame_idx = 0;  // current_time_ = t0_ + frame_idx*control_dt_  (ENV:557,631,786); table mode: reset() takes the random start row out of t0_
    int ref_row() const { int r = frame_idx < 0 ? 0 : frame_idx; return r > ref_rows - 1 ? ref_rows - 1 : r; }   // clamped like the CUDA path (the reference reads unchecked)
    // ---- Crutial: True -- meteor spheres (ENV:273-283, 608-611, 717-741, 815-861).  The reference drops CubeNum steel spheres from
    // 1 m above the trunk every 5 gait periods; cube_place_radius is 0 (ENV:1976), so they coincide and are modelled as ONE
    // sphere of CubeNum times the mass (identical frictionless bodies hitting the same point).  New specification (RaiSim is
    // not available): sphere <-> trunk box and sphere <-> ground only, handled once per control step before the physics substeps.
    int num_cube = 0;
    T met_p[3] = {0, 0, 0}, met_v[3] = {0, 0, 0}, met_r = 0, met_m = 0; int met_mode = 0;   // mode 0 = static (just placed), 1 = falling
    T jointRef_[NJ], jointRefLast_[NJ], jointDotRef_[NJ], EndEffectorRef_[NJ], EndEffector_[NJ], EndEffectorOffset_[NJ];
    T command[3] = {0, 0, 0}, command_filtered[3] = {0, 0, 0};
    T contact_[4] = {0, 0, 0, 0}, contact_filtered[4] = {0, 0, 0, 0};
    T contact_force_norm[4] = {0, 0, 0, 0}, contact_vel_norm[4] = {0, 0, 0, 0};
    T EndEffectorReward = 0, BodyCenterReward = 0, BodyAttitudeReward = 0, JointReward = 0, JointDotReward = 0,
      VelocityReward = 0, TorqueReward = 0, ContactReward = 0;
    T filter_para = 0;
    unsigned long itera = 0;
    // contact bookkeeping of the last integrate()
    int n_contacts = 0; int contact_kind[MAXC]; T contact_impulse[MAXC][3];
    int foot_in_contact[4] = {0, 0, 0, 0}; T foot_impulse[4][3];
    int last_solver_sweeps = 0;
    // distance of this control step's discrete decisions from their thresholds (test infrastructure: fp32 and fp64 can only agree on
    // a decision whose margin exceeds fp32 resolution).  geo: min |signed gap| of any toe / trunk corner over the substeps [m];
    // rest: min |v_n + threshold| over new contacts [m/s]; term: min distance of the termination tests (ENV:1560) from their bounds
    // cone: min relative distance |lambda_t| vs mu lambda_n of any single-contact visit from the stick / slide boundary (the one-step
    // sliding rule is not continuous across it: the stick impulse points along d, the sliding one along G_tt d)
    mutable T margin_geo = T(1e30), margin_rest = T(1e30), margin_term = T(1e30), margin_cone = T(1e30);
    // optional reference table (ManualTraj False; ENV:17-21, VEC:158-182)
    const float* ref = nullptr; int ref_rows = 0, frame_max = 0, frame_len = 0;
    Terrain terrain;               // valid() only when Terrain: True
    T ground_height() const { if (!terrain.valid()) return T(0); T hh; V3<T> nn; terrain.sample(gc_[0], gc_[1], hh, nn); return hh; }

    T current_time() const { return t0_ + T(frame_idx) * control_dt_; }

    // ---------------------------------------------------------------- ctor ENV:216-532
    void configure(const Cfg& c, uint32_t env_id_, uint32_t seed_) {
        env_id = env_id_; seed = seed_;
        abad_ = T(c.get("abad")); period_ = T(c.get("period")); lam_ = T(c.get("lam")); stand_height_ = T(c.get("stand_height"));
        up_height_ = T(c.get("up_height")); up_height_max_ = up_height_; down_height_ = T(c.get("down_height")); gait_step_ = T(c.get("gait_step"));
        Vx_max = T(c.get("Vx")); Vy_max = T(c.get("Vy")); Vy_min = -Vy_max; omega_max = T(c.get("Omega")); omega_min = -omega_max;  // ENV:1606-1611
        Lean_middle_front = T(c.get("LeanFront")); Lean_middle_hind = T(c.get("LeanHind"));
        flag_terrain = c.flag("Terrain"); flag_manual = c.flag("Manual"); flag_crucial = c.flag("Crutial"); flag_filter = c.flag("Filter");
        (void)c.get("Camera"); flag_StochasticDynamics = c.flag("StochasticDynamics"); flag_HeightVariable = c.flag("HeightVariable");
        flag_TimeBasedContact = c.flag("TimeBasedContact"); flag_ManualTraj = c.flag("ManualTraj"); flag_MotorDynamics = c.flag("MotorDynamics");
        flag_ObsFilter = c.flag("ObsFilter"); flag_WildCat = c.flag("WILDCAT"); flag_ForceDisturbance = c.flag("ForceDisturbance"); (void)c.get("Convert2Torque");
        terminalRewardCoeff_ = T(c.get("terminalRewardCoeff")); EECoeff = T(c.get("EndEffectorRewardCoeff")); BodyPosCoeff = T(c.get("BodyPosRewardCoeff"));
        BodyAttiCoeff = T(c.get("BodyAttitudeRewardCoeff")); JointMimicCoeff = T(c.get("JointRewardCoeff")); VelKeepCoeff = T(c.get("VelRewardCoeff"));
        TorqueCoeff = T(c.get("TorqueCoeff")); ContactCoeff = T(c.get("ContactCoeff"));
        stiffness = T(c.get("Stiffness")); (void)c.get("Stiffness_Low"); abad_ratio = T(c.get("AbadRatio")); damping = T(c.get("Damping"));
        freq = T(c.get("Freq")); max_time = T(c.get("max_time")); num_cube = int(c.get("CubeNum")); (void)c.get("FPS");
        actionNoise = T(c.get("ActionNoise")); noise_flag = T(c.get("ObsNoise")); gaitType = (int)c.get("GaitType");
        MotorMaxTorque = T(c.get("MotorMaxTorque")); MotorCriticalSpeed = T(c.get("MotorCriticalSpeed")); MotorMaxSpeed = T(c.get("MotorMaxSpeed"));
        simulation_dt_ = T(c.get("simulation_dt")); control_dt_ = T(c.get("control_dt"));   // VEC:151-152
        // solver / model switches (new-spec, optional)
        model.joint_damping = T(c.get_or("joint_damping", double(model.joint_damping)));
        solver_iters = (int)c.get_or("solver_iters", 30); solver_tol = T(c.get_or("solver_tol", 1e-5)); slide_iters = (int)c.get_or("slide_iters", 1); solver_jacobi = (int)c.get_or("solver_jacobi", 1); jacobi_sweeps = (int)c.get_or("jacobi_sweeps", 10); slide_exact = (int)c.get_or("slide_exact", 0);
        mu = T(c.get_or("friction", 0.6)); restitution = T(c.get_or("restitution", 0.2)); rest_threshold = T(c.get_or("restitution_threshold", 0.01));

        // gc_init_ ENV:317-322
        const T init[NQ] = {0, 0, T(0.35), 1, 0, 0, 0, -abad_, T(-0.78), T(1.57), abad_, T(-0.78), T(1.57), -abad_, T(-0.78), T(1.57), abad_, T(-0.78), T(1.57)};
        for (int i = 0; i < NQ; ++i) gc_init_[i] = gc_[i] = init[i];
        for (int i = 0; i < NV; ++i) { gv_[i] = 0; gf_[i] = 0; }
        const T eeo[NJ] = {T(0.19), T(-0.058), 0, T(0.19), T(0.058), 0, T(-0.19), T(-0.058), 0, T(-0.19), T(0.058), 0};  // ENV:331-334
        for (int i = 0; i < NJ; ++i) {
            EndEffectorOffset_[i] = eeo[i]; EndEffectorRef_[i] = 0; EndEffector_[i] = 0;
            bool is_abad = (i % 3 == 0);
            jointPgain[i] = is_abad ? stiffness * abad_ratio : stiffness;   // ENV:340-343
            jointDgain[i] = is_abad ? damping * abad_ratio : damping;       // ENV:347-350
            torque[i] = torque_last[i] = 0; pTarget12_[i] = pTarget12Last_[i] = 0;
            torque_limit[i] = (i % 3 == 2) ? T(27) : T(18);                  // ENV:354
            actionMean_[i] = gc_init_[7 + i];                               // ENV:371
            jointRefLast_[i] = 0; jointDotRef_[i] = 0;
        }
        // obMean_ / obStd_ ENV:375-393
        for (int i = 0; i < 35; ++i) { obMean_[i] = 0; obStd_[i] = 1; obDouble_[i] = 0; obDouble_last_[i] = 0; }
        obMean_[0] = (Vx_max + Vx_min) / 2; obMean_[1] = (Vy_max + Vy_min) / 2; obMean_[2] = (omega_max + omega_min) / 2;
        for (int i = 0; i < 12; ++i) obMean_[5 + i] = gc_init_[7 + i];
        obMean_[31] = 1;
        const T vstd[3] = {5, 35, 40};
        for (int i = 0; i < 12; ++i) obStd_[17 + i] = vstd[i % 3];
        for (int i = 0; i < 3; ++i) { obStd_[29 + i] = T(0.7); obStd_[32 + i] = T(3.0); }
        max_len = mth::sqrt(l_hip_ * l_hip_ + (l_calf_ + l_thigh_) * (l_calf_ + l_thigh_));   // ENV:395
        filter_para = flag_filter ? (1 - freq * control_dt_) : 0;                              // ENV:396
        switch (gaitType) {                                                                    // ENV:398-409
            case 0: phase_[0] = T(0.5); phase_[1] = 0; phase_[2] = 0; phase_[3] = T(0.5); break;
            case 1: phase_[0] = T(0.5); phase_[1] = T(0.5); phase_[2] = 0; phase_[3] = 0; break;
            case 2: phase_[0] = 0; phase_[1] = T(0.25); phase_[2] = T(0.5); phase_[3] = T(0.75); break;
            default: phase_[0] = phase_[1] = phase_[2] = phase_[3] = 0;
        }
        for (int l = 0; l < 4; ++l) { jointRef_[3 * l] = (l % 2 == 0) ? -abad_ : abad_; jointRef_[3 * l + 1] = 0; jointRef_[3 * l + 2] = 0; }  // ENV:415-418
        if (flag_ObsFilter) ObsFilterAlpha = T(2.0 * 3.14) * control_dt_ * ObsFilterFreq / (T(2.0 * 3.14) * control_dt_ * ObsFilterFreq + T(1.0));  // ENV:425-426
        if (flag_StochasticDynamics) randomize_dynamics();   // ENV:435-477
    }

    // ENV:435-477.  Distributions kept; draws come from Philox (tick = 0xFFFFFFFF).
    void randomize_dynamics() {
        uint32_t r[4];
        Philox::gen(seed, env_id, 0xFFFFFFFFu, P_DR_MATERIAL, r);
        mu = T(u01(r[0])) * T(0.6) + T(0.4); restitution = T(u01(r[1])) * T(0.3); rest_threshold = T(u01(r[2])) * T(2.0);   // ENV:440-442
        for (int i = 0; i < NB; ++i) {
            Philox::gen(seed, env_id, 0xFFFFFFFFu, P_DR_MASS + i, r);
            T k = (T(u01(r[0])) - T(0.5)) / T(0.5) * mass_distrubance_ratio + T(1.0);       // ENV:454
            model.b[i].mass = model.b[i].mass * k;                                         // ENV:456 (inertia tensors untouched)
            Philox::gen(seed, env_id, 0xFFFFFFFFu, P_DR_COM + i, r);
            for (int a = 0; a < 3; ++a) model.b[i].com[a] += T(usym(r[a])) * com_distrubance;   // ENV:463-465
        }
        Philox::gen(seed, env_id, 0xFFFFFFFFu, P_DR_CALF, r);
        T dz = (T(u01(r[0])) - T(0.5)) / T(0.5) * calf_distrubance;                         // ENV:472
        for (int l = 0; l < 4; ++l) model.b[3 + 3 * l].off.z += dz;                         // ENV:473-476 (bodies 3,6,9,12 = shanks)
    }

    // ================================================================ rigid-body dynamics (RaiSim replacement)
    struct Kin {
        M3<T> R[NB];      // world rotation of each body
        V3<T> pj[NB];     // joint origin, relative to the trunk origin, world axes
        V3<T> pc[NB];     // COM, relative to the trunk origin
        V3<T> ax[NB];     // joint axis in world
        V3<T> w[NB];      // angular velocity
        V3<T> vj[NB];     // velocity of joint origin
        V3<T> vc[NB];     // velocity of COM
        V3<T> toe[4];     // toe frame origins (relative to trunk origin)
        V3<T> vtoe[4];
    };
    void kinematics(const T* gc, const T* gv, Kin& k) const {
        k.R[0] = quat_to_rot(gc + 3);
        k.pj[0] = V3<T>(0, 0, 0); k.pc[0] = k.R[0] * model.b[0].com; k.ax[0] = V3<T>(0, 0, 0);
        k.w[0] = V3<T>(gv[3], gv[4], gv[5]); k.vj[0] = V3<T>(gv[0], gv[1], gv[2]);
        k.vc[0] = k.vj[0] + cross(k.w[0], k.pc[0]);
        for (int i = 1; i < NB; ++i) {
            const Body<T>& b = model.b[i]; int p = b.parent;
            k.R[i] = k.R[p] * axis_angle(b.axis, gc[7 + (i - 1)]);
            k.pj[i] = k.pj[p] + k.R[p] * b.off;
            k.ax[i] = k.R[p] * b.axis;
            k.pc[i] = k.pj[i] + k.R[i] * b.com;
            k.vj[i] = k.vj[p] + cross(k.w[p], k.pj[i] - k.pj[p]);
            k.w[i] = k.w[p] + gv[6 + (i - 1)] * k.ax[i];
            k.vc[i] = k.vj[i] + cross(k.w[i], k.pc[i] - k.pj[i]);
        }
        for (int l = 0; l < 4; ++l) {
            int s = 3 + 3 * l;
            k.toe[l] = k.pj[s] + k.R[s] * model.toe_off;
            k.vtoe[l] = k.vj[s] + cross(k.w[s], k.toe[l] - k.pj[s]);
        }
    }
    // is body `a` an ancestor of (or equal to) body `i`?
    bool on_path(int a, int i) const { while (i >= 0) { if (i == a) return true; i = model.b[i].parent; } return false; }
    // point Jacobian (3x18) of a point x (relative to trunk origin) fixed to body i
    void point_jacobian(const Kin& k, int body, const V3<T>& x, T J[3][NV]) const {
        for (int r = 0; r < 3; ++r) for (int c = 0; c < NV; ++c) J[r][c] = 0;
        for (int r = 0; r < 3; ++r) J[r][r] = 1;
        // d(omega x x)/d omega = -[x]x
        J[0][4] = x.z;  J[0][5] = -x.y; J[1][3] = -x.z; J[1][5] = x.x; J[2][3] = x.y; J[2][4] = -x.x;
        for (int j = 1; j < NB; ++j) if (on_path(j, body)) {
            V3<T> c = cross(k.ax[j], x - k.pj[j]);
            J[0][6 + j - 1] = c.x; J[1][6 + j - 1] = c.y; J[2][6 + j - 1] = c.z;
        }
    }
    // Mass matrix by explicit Jacobians: M = sum_i m_i Jv_i^T Jv_i + Jw_i^T (R I R^T) Jw_i + diag(rotor)
    void mass_matrix(const Kin& k, T M[NV][NV]) const {
        for (int a = 0; a < NV; ++a) for (int b = 0; b < NV; ++b) M[a][b] = 0;
        for (int i = 0; i < NB; ++i) {
            T Jv[3][NV]; point_jacobian(k, i, k.pc[i], Jv);
            T Jw[3][NV]; for (int r = 0; r < 3; ++r) for (int c = 0; c < NV; ++c) Jw[r][c] = 0;
            for (int r = 0; r < 3; ++r) Jw[r][3 + r] = 1;
            for (int j = 1; j < NB; ++j) if (on_path(j, i)) { Jw[0][6 + j - 1] = k.ax[j].x; Jw[1][6 + j - 1] = k.ax[j].y; Jw[2][6 + j - 1] = k.ax[j].z; }
            M3<T> Iw = k.R[i] * model.b[i].I * transpose(k.R[i]);
            T IJ[3][NV];
            for (int r = 0; r < 3; ++r) for (int c = 0; c < NV; ++c) { T s = 0; for (int q = 0; q < 3; ++q) s += Iw.m[r][q] * Jw[q][c]; IJ[r][c] = s; }
            for (int a = 0; a < NV; ++a) for (int b = 0; b < NV; ++b) {
                T s = 0; for (int r = 0; r < 3; ++r) s += model.b[i].mass * Jv[r][a] * Jv[r][b] + Jw[r][a] * IJ[r][b];
                M[a][b] += s;
            }
        }
        for (int j = 1; j < NB; ++j) M[6 + j - 1][6 + j - 1] += model.b[j].rotor;
    }
    // Nonlinear term h(q,u) (Coriolis/centrifugal + gravity, M u' + h = tau + J^T f), classical
    // Newton-Euler bias accelerations projected through the body Jacobians.
    void nonlinearities(const Kin& k, const T* gv, T h[NV]) const {
        V3<T> al[NB], aj[NB];
        for (int a = 0; a < NV; ++a) h[a] = 0;
        al[0] = V3<T>(0, 0, 0); aj[0] = V3<T>(0, 0, 0);
        for (int i = 0; i < NB; ++i) {
            const Body<T>& b = model.b[i]; int p = b.parent;
            if (i > 0) {
                V3<T> d = k.pj[i] - k.pj[p];
                aj[i] = aj[p] + cross(al[p], d) + cross(k.w[p], cross(k.w[p], d));
                al[i] = al[p] + gv[6 + i - 1] * cross(k.w[p], k.ax[i]);
            }
            V3<T> r = k.pc[i] - k.pj[i];
            V3<T> ac = aj[i] + cross(al[i], r) + cross(k.w[i], cross(k.w[i], r));
            V3<T> F = b.mass * (ac + V3<T>(0, 0, model.gravity));
            M3<T> Iw = k.R[i] * b.I * transpose(k.R[i]);
            V3<T> N = Iw * al[i] + cross(k.w[i], Iw * k.w[i]);
            T Jv[3][NV]; point_jacobian(k, i, k.pc[i], Jv);
            for (int a = 0; a < NV; ++a) h[a] += Jv[0][a] * F.x + Jv[1][a] * F.y + Jv[2][a] * F.z;
            for (int r3 = 0; r3 < 3; ++r3) h[3 + r3] += N[r3];
            for (int j = 1; j < NB; ++j) if (on_path(j, i)) h[6 + j - 1] += dot(k.ax[j], N);
        }
    }

    struct Contact { int kind; /*0..3 foot, 4.. box corner*/ int body; V3<T> point; V3<T> n; };
    // collision detection: toe spheres and trunk box corners against the plane z = 0 (ENV:268 addGround) or the heightfield
    // (ENV:264 addHeightMap): signed distance to the local triangle plane, contact point = centre - r n
    int detect_contacts(const T* gc, const Kin& k, Contact* cs) const {
        int n = 0;
        for (int l = 0; l < 4; ++l) {
            T hh = 0; V3<T> nn(0, 0, 1);
            if (terrain.valid()) terrain.sample(gc[0] + k.toe[l].x, gc[1] + k.toe[l].y, hh, nn);
            T gap = (gc[2] + k.toe[l].z - hh) * nn.z - model.toe_radius;
            if (mth::fabs(gap) < margin_geo) margin_geo = mth::fabs(gap);
            if (gap <= T(0)) {
                cs[n].kind = l; cs[n].body = 3 + 3 * l; cs[n].n = nn;
                cs[n].point = k.toe[l] - model.toe_radius * nn; ++n;
            }
        }
        int nbox = 0;
        for (int c = 0; c < 8 && nbox < 4; ++c) {
            V3<T> loc((c & 1) ? model.box_half.x : -model.box_half.x, (c & 2) ? model.box_half.y : -model.box_half.y, (c & 4) ? model.box_half.z : -model.box_half.z);
            V3<T> w = k.R[0] * loc;
            T hh = 0; V3<T> nn(0, 0, 1);
            if (terrain.valid()) terrain.sample(gc[0] + w.x, gc[1] + w.y, hh, nn);
            if (mth::fabs((gc[2] + w.z - hh) * nn.z) < margin_geo) margin_geo = mth::fabs((gc[2] + w.z - hh) * nn.z);
            if ((gc[2] + w.z - hh) * nn.z <= T(0)) { cs[n].kind = 4 + c; cs[n].body = 0; cs[n].n = nn; cs[n].point = w; ++n; ++nbox; }
        }
        return n;
    }
    // contact frame rows (t1, t2, n): t1 = x-axis projected on the tangent plane, t2 = n x t1 (identity on flat ground)
    static void contact_frame(const V3<T>& n, V3<T> D[3]) {
        V3<T> t1(T(1) - n.x * n.x, -n.x * n.y, -n.x * n.z);
        T inv = T(1) / mth::sqrt(dot(t1, t1)); t1 = inv * t1;
        D[0] = t1; D[1] = cross(n, t1); D[2] = n;
    }

    // One world_->integrate() (ENV:768).  tau = joint torques (12).
    void integrate(const T* tau12) {
        const T dt = simulation_dt_;
        Kin k; kinematics(gc_, gv_, k);
        static thread_local T M[NV][NV]; T h[NV];
        mass_matrix(k, M); nonlinearities(k, gv_, h);
        Chol<T, NV> ch; if (!ch.factor(M)) throw std::runtime_error("mass matrix not PD");
        T rhs[NV], ufree[NV];
        for (int a = 0; a < 6; ++a) { gf_[a] = 0; rhs[a] = -h[a]; }
        for (int j = 0; j < NJ; ++j) { gf_[6 + j] = tau12[j]; rhs[6 + j] = tau12[j] - model.joint_damping * gv_[6 + j] - h[6 + j]; }
        ch.solve(rhs, ufree);
        for (int a = 0; a < NV; ++a) ufree[a] = gv_[a] + dt * ufree[a];

        Contact cs[MAXC]; int nc = detect_contacts(gc_, k, cs);
        n_contacts = nc;
        for (int l = 0; l < 4; ++l) { foot_in_contact[l] = 0; foot_impulse[l][0] = foot_impulse[l][1] = foot_impulse[l][2] = 0; }
        T unew[NV]; for (int a = 0; a < NV; ++a) unew[a] = ufree[a];
        last_solver_sweeps = 0;
        if (nc > 0) {
            static thread_local T J[MAXC][3][NV], W[MAXC][3][NV];   // W = (M^-1 J^T)^T
            T G[3 * MAXC][3 * MAXC], c[3 * MAXC], vtarget[MAXC], lam[3 * MAXC];
            for (int i = 0; i < nc; ++i) {
                point_jacobian(k, cs[i].body, cs[i].point, J[i]);
                if (terrain.valid()) {   // express the three rows in the contact frame (t1, t2, n)
                    V3<T> D[3]; contact_frame(cs[i].n, D);
                    for (int a = 0; a < NV; ++a) { V3<T> col(J[i][0][a], J[i][1][a], J[i][2][a]); for (int r = 0; r < 3; ++r) J[i][r][a] = dot(D[r], col); }
                }
                for (int r = 0; r < 3; ++r) ch.solve(J[i][r], W[i][r]);
            }
            for (int i = 0; i < nc; ++i) for (int r = 0; r < 3; ++r) {
                for (int j = 0; j < nc; ++j) for (int s = 0; s < 3; ++s) { T acc = 0; for (int a = 0; a < NV; ++a) acc += J[i][r][a] * W[j][s][a]; G[3 * i + r][3 * j + s] = acc; }
                T acc = 0, pre = 0; for (int a = 0; a < NV; ++a) { acc += J[i][r][a] * ufree[a]; pre += J[i][r][a] * gv_[a]; }
                c[3 * i + r] = acc; lam[3 * i + r] = 0;
                if (r == 2 && mth::fabs(pre + rest_threshold) < margin_rest) margin_rest = mth::fabs(pre + rest_threshold);
                if (r == 2) vtarget[i] = (pre < -rest_threshold) ? -restitution * pre : T(0);   // Newton restitution above the threshold speed
            }
            // per-contact Gauss-Seidel (normal = +z on the plane, so contact frame == world frame)
            for (int sweep = 0; sweep < solver_iters; ++sweep) {
                T maxd = 0, maxl = 0;
                // schedule: the (<= 4) foot contacts are updated simultaneously from the impulses of the previous sweep
                // (block Jacobi: feet couple only weakly, through the trunk), the trunk-box corners one after the other
                // with the freshest impulses (Gauss-Seidel: they sit on the same rigid body).  solver_jacobi = 0 gives
                // plain Gauss-Seidel over all contacts.
                T lam_prev[3 * MAXC]; for (int q = 0; q < 3 * nc; ++q) lam_prev[q] = lam[q];
                for (int i = 0; i < nc; ++i) {
                    const T* lsrc = (solver_jacobi && sweep < jacobi_sweeps && cs[i].kind < 4) ? lam_prev : lam;
                    T v[3]; for (int r = 0; r < 3; ++r) { T acc = c[3 * i + r]; for (int q = 0; q < 3 * nc; ++q) acc += G[3 * i + r][q] * lsrc[q]; v[r] = acc; }
                    T Gii[3][3], Ginv[3][3]; for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) Gii[r][s] = G[3 * i + r][3 * i + s];
                    inv3(Gii, Ginv);
                    T lo[3] = {lam[3 * i], lam[3 * i + 1], lam[3 * i + 2]}, ln[3];
                    solve_one_contact(v, Gii, Ginv, lo, vtarget[i], ln);
                    for (int r = 0; r < 3; ++r) { T d = mth::fabs(ln[r] - lo[r]); if (d > maxd) maxd = d; if (mth::fabs(ln[r]) > maxl) maxl = mth::fabs(ln[r]); lam[3 * i + r] = ln[r]; }
                }
                last_solver_sweeps = sweep + 1;
                if (maxd <= solver_tol * maxl) break;
            }
            for (int i = 0; i < nc; ++i) {
                for (int r = 0; r < 3; ++r) for (int a = 0; a < NV; ++a) unew[a] += W[i][r][a] * lam[3 * i + r];
                contact_kind[i] = cs[i].kind; for (int r = 0; r < 3; ++r) contact_impulse[i][r] = lam[3 * i + r];
                if (cs[i].kind < 4) { foot_in_contact[cs[i].kind] = 1; for (int r = 0; r < 3; ++r) foot_impulse[cs[i].kind][r] = lam[3 * i + r]; }
            }
        }
        // semi-implicit Euler: positions advance with the new velocity
        for (int a = 0; a < NV; ++a) gv_[a] = unew[a];
        for (int a = 0; a < 3; ++a) gc_[a] += dt * gv_[a];
        {   // orientation: rotate by |w| dt about w (world frame): q+ = dq (x) q
            T wx = gv_[3], wy = gv_[4], wz = gv_[5];
            T wn = mth::sqrt(wx * wx + wy * wy + wz * wz), th = wn * dt;
            T kk, cw;
            if (th > T(1e-8)) { kk = mth::sin(th / 2) / wn; cw = mth::cos(th / 2); } else { kk = dt / 2; cw = T(1); }
            T dq[4] = {cw, kk * wx, kk * wy, kk * wz}, q[4] = {gc_[3], gc_[4], gc_[5], gc_[6]}, o[4];
            o[0] = dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2] - dq[3] * q[3];
            o[1] = dq[0] * q[1] + dq[1] * q[0] + dq[2] * q[3] - dq[3] * q[2];
            o[2] = dq[0] * q[2] - dq[1] * q[3] + dq[2] * q[0] + dq[3] * q[1];
            o[3] = dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1] + dq[3] * q[0];
            T nn = T(1) / mth::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
            for (int a = 0; a < 4; ++a) gc_[3 + a] = o[a] * nn;
        }
        for (int j = 0; j < NJ; ++j) gc_[7 + j] += dt * gv_[6 + j];
    }

    // Single-contact solve with the other contacts frozen.  v = current contact velocity (with the current
    // impulse lo applied), G = 3x3 Delassus block, n = +z.  Hard contact: either separation (lambda = 0),
    // stick (v_t = 0, v_n = vt_n, impulse inside the cone) or slide (impulse on the cone boundary,
    // opposing the sliding velocity, v_n = vt_n), the latter by a fixed-point on the sliding direction.
    void solve_one_contact(const T v[3], const T G[3][3], const T Ginv[3][3], const T lo[3], T vtn, T ln[3]) const {
        T e[3] = {v[0], v[1], v[2] - vtn};
        T ls[3]; for (int r = 0; r < 3; ++r) ls[r] = lo[r] - (Ginv[r][0] * e[0] + Ginv[r][1] * e[1] + Ginv[r][2] * e[2]);
        if (!(ls[2] > T(0))) { ln[0] = ln[1] = ln[2] = 0; return; }
        T lt = mth::sqrt(ls[0] * ls[0] + ls[1] * ls[1]);
        { T mc = mth::fabs(lt - mu * ls[2]) / (mu * ls[2]); if (mc < margin_cone) margin_cone = mc; }
        if (lt <= mu * ls[2]) { ln[0] = ls[0]; ln[1] = ls[1]; ln[2] = ls[2]; return; }
        // b = velocity with zero impulse at this contact
        T b[3]; for (int r = 0; r < 3; ++r) b[r] = v[r] - (G[r][0] * lo[0] + G[r][1] * lo[1] + G[r][2] * lo[2]);
        T dx = ls[0] / lt, dy = ls[1] / lt;
        if (slide_exact) {
            // impulse on the cone boundary lambda = ln (mu d, 1), d = (cos th, sin th): v_n(lambda) = vtn fixes ln(th); the sliding
            // velocity v_t(th) must be anti-parallel to d (maximal dissipation).  f(th) = d x v_t changes sign at the solution;
            // scan the circle from the stick direction, bisect the dissipative root closest to it.
            auto eval = [&](T th, T& lnz_, T& f, T& dotp) {
                T cx = mth::cos(th), sy = mth::sin(th);
                T den = G[2][2] + mu * (G[2][0] * cx + G[2][1] * sy);
                lnz_ = (den > T(1e-12)) ? (vtn - b[2]) / den : T(0); if (lnz_ < T(0)) lnz_ = 0;
                T l0 = mu * lnz_ * cx, l1 = mu * lnz_ * sy;
                T vx = b[0] + G[0][0] * l0 + G[0][1] * l1 + G[0][2] * lnz_, vy = b[1] + G[1][0] * l0 + G[1][1] * l1 + G[1][2] * lnz_;
                f = cx * vy - sy * vx; dotp = cx * vx + sy * vy;
            };
            const T th0 = mth::atan2(dy, dx); const int NS = 64; const T step = T(2 * BP5O_PI) / NS;
            T best_lo = 0, best_hi = 0; bool found = false; T best_dist = T(1e30);
            T fp, lp, dp; eval(th0, lp, fp, dp);
            for (int s2 = 1; s2 <= NS; ++s2) {
                // alternate +/- around th0 so the closest bracket is found first
                for (int sgn = -1; sgn <= 1; sgn += 2) {
                    T a = th0 + T(sgn) * step * T(s2 - 1), bnd = th0 + T(sgn) * step * T(s2);
                    T fa, la, da, fb, lb, db; eval(a, la, fa, da); eval(bnd, lb, fb, db);
                    if ((fa <= T(0)) != (fb <= T(0)) && (da < T(0) || db < T(0))) {
                        T dist = step * T(s2 - 1);
                        if (dist < best_dist) { best_dist = dist; best_lo = a; best_hi = bnd; found = true; }
                    }
                }
                if (found) break;
            }
            if (found) {
                T a = best_lo, bnd = best_hi, fa, la, da; eval(a, la, fa, da);
                for (int it = 0; it < 60; ++it) { T m = (a + bnd) / 2, fm, lm, dm; eval(m, lm, fm, dm); if ((fm <= T(0)) == (fa <= T(0))) { a = m; fa = fm; } else bnd = m; }
                T th = (a + bnd) / 2, lz, f, dp2; eval(th, lz, f, dp2);
                ln[0] = mu * lz * mth::cos(th); ln[1] = mu * lz * mth::sin(th); ln[2] = lz; return;
            }
        }
        T lnz = 0;
        for (int it = 0; it < slide_iters; ++it) {
            T den = G[2][2] + mu * (G[2][0] * dx + G[2][1] * dy);
            if (!(den > T(1e-12))) break;
            lnz = (vtn - b[2]) / den; if (lnz < T(0)) lnz = 0;
            T l0 = mu * lnz * dx, l1 = mu * lnz * dy;
            T vx = b[0] + G[0][0] * l0 + G[0][1] * l1 + G[0][2] * lnz;
            T vy = b[1] + G[1][0] * l0 + G[1][1] * l1 + G[1][2] * lnz;
            T vn = mth::sqrt(vx * vx + vy * vy);
            if (vn > T(1e-9)) { dx = -vx / vn; dy = -vy / vn; }
        }
        {   // final normal solve with the converged direction
            T den = G[2][2] + mu * (G[2][0] * dx + G[2][1] * dy);
            if (den > T(1e-12)) { lnz = (vtn - b[2]) / den; if (lnz < T(0)) lnz = 0; }
        }
        ln[0] = mu * lnz * dx; ln[1] = mu * lnz * dy; ln[2] = lnz;
    }

    // ================================================================ env logic (ENV line by line)
    // ENV:1273-1312
    void torque_clamp(const T* gv_temp) {
        T r = MotorMaxTorque / (MotorMaxSpeed - MotorCriticalSpeed);
        for (int i = 0; i < NJ; ++i) {
            T ratio = ((i + 1) % 3 == 0) ? T(1.55f) : T(1.0f);
            T s = gv_temp[i + 6] * ratio;
            T up = (s > MotorCriticalSpeed) ? (MotorMaxTorque - (s - MotorCriticalSpeed) * r) : MotorMaxTorque;
            up = up * ratio;
            T low = (s < -MotorCriticalSpeed) ? ((-MotorMaxSpeed - s) / (-MotorMaxSpeed + MotorCriticalSpeed) * -MotorMaxTorque) : -MotorMaxTorque;
            low = low * ratio;
            torque[i] = mth::fmax(mth::fmin(torque[i], up), low);
        }
    }

    // ENV:1687-1751
    void inverse_kinematics(T x, T y, T z, T l_hip, T l_thigh, T l_calf, T* theta, bool is_right) const {
        T ll = mth::sqrt(x * x + y * y + z * z);
        if (ll > max_len) { x = x * (max_len - T(1e-5)) / ll; y = y * (max_len - T(1e-5)) / ll; z = z * (max_len - T(1e-5)) / ll; }
        T temp, temp1, temp2 = 0;
        if (is_right) { temp = (-z * l_hip - mth::sqrt(y * y * (z * z + y * y - l_hip * l_hip))) / (z * z + y * y); if (mth::fabs(temp) <= 1) theta[0] = mth::asin(temp); }
        else          { temp = ( z * l_hip + mth::sqrt(y * y * (z * z + y * y - l_hip * l_hip))) / (z * z + y * y); if (mth::fabs(temp) <= 1) theta[0] = mth::asin(temp); }
        T lr = mth::sqrt(x * x + y * y + z * z - l_hip * l_hip);
        lr = (lr > (l_thigh + l_calf)) ? (l_thigh + l_calf - T(1e-4)) : lr;
        temp = (l_thigh * l_thigh + l_calf * l_calf - lr * lr) / 2 / l_thigh / l_calf + T(1e-5);
        if (mth::fabs(temp) <= 1) theta[2] = -(T(BP5O_PI) - mth::acos(temp));
        temp1 = x / lr;
        temp2 = (lr * lr + l_thigh * l_thigh - l_calf * l_calf) / 2 / lr / l_thigh - T(1e-5);
        if (mth::fabs(temp1) <= 1 && mth::fabs(temp2) <= 1) theta[1] = mth::acos(temp2) - mth::asin(temp1);
    }

    // toe target for leg i at absolute time tt   (body of the loops at ENV:1802-1842 / ENV:1844-1885)
    void leg_reference(int i, T tt, T* joint3, T* toe3) const {
        T real_phase = mth::fmod(tt + phase_[i] * period_, period_) / period_;
        // NB: the reference computes fmod(current_time_ + phase*period [- dt], period); tt already carries the -dt
        T anti_flag = (i < 2) ? T(1.0) : T(-1.0);
        V3<T> p0, pf, toe;
        if (real_phase < lam_) {
            T temp_r = real_phase / lam_;
            p0 = V3<T>(gait_step_ / T(2.0), side_step_ / T(2.0) + anti_flag * rot_step_ / T(2.0), -stand_height_);
            pf = V3<T>(-gait_step_ / T(2.0), -side_step_ / T(2.0) + -anti_flag * rot_step_ / T(2.0), -stand_height_);
            toe = cubicBezier(p0, pf, temp_r);
        } else {
            T temp_r = (real_phase - lam_) / (T(1.0) - lam_);
            pf = V3<T>(gait_step_ / T(2.0), side_step_ / T(2.0) + anti_flag * rot_step_ / T(2.0), -stand_height_);
            p0 = V3<T>(-gait_step_ / T(2.0), -side_step_ / T(2.0) + -anti_flag * rot_step_ / T(2.0), -stand_height_);
            toe = Bezier2(p0, pf, temp_r, up_height_);
        }
        T temp_offset[4] = {-l_hip_ + Lean_middle_front, l_hip_ - Lean_middle_front, -l_hip_ + Lean_middle_hind, l_hip_ - Lean_middle_hind};  // ENV:1795-1798
        T th[3] = {0, 0, 0};
        inverse_kinematics(toe.x, toe.y + temp_offset[i], toe.z, l_hip_, l_thigh_, l_calf_, th, i == 0 || i == 2);
        joint3[0] = th[0]; joint3[1] = -th[1]; joint3[2] = -th[2];                                             // ENV:1879-1881
        toe3[0] = toe.x; toe3[1] = toe.y; toe3[2] = toe.z;
    }

    // ENV:1756-1890
    void gait_generator_manual(bool is_first) {
        gait_step_ = command_filtered[0] * lam_ * period_;            // ENV:1772
        gait_step_ = flag_WildCat ? -gait_step_ : gait_step_;         // ENV:1773
        side_step_ = command_filtered[1] * lam_ * period_;            // ENV:1775
        rot_step_ = command_filtered[2] * period_ * T(0.4);           // ENV:1777
        if (flag_HeightVariable) {                                    // ENV:1779-1792
            T ratio = mth::fabs(command_filtered[0]) / Vx_max;
            if (Vy_max > 0) ratio = mth::fmax(ratio, mth::fabs(command_filtered[1]) / Vy_max);
            if (omega_max > 0) ratio = mth::fmax(ratio, mth::fabs(command_filtered[2] / omega_max));
            up_height_ = (ratio > T(0.1)) ? up_height_max_ : ratio * up_height_max_;
        }
        T toe3[3];
        if (is_first) for (int i = 0; i < 4; ++i) leg_reference(i, current_time() - control_dt_, jointRefLast_ + 3 * i, toe3);   // ENV:1799-1843
        for (int i = 0; i < 4; ++i) {
            leg_reference(i, current_time(), jointRef_ + 3 * i, toe3);
            for (int a = 0; a < 3; ++a) EndEffectorRef_[3 * i + a] = toe3[a] + EndEffectorOffset_[3 * i + a];   // ENV:1882-1889
        }
        for (int j = 0; j < NJ; ++j) { jointDotRef_[j] = (jointRef_[j] - jointRefLast_[j]) / control_dt_; jointRefLast_[j] = jointRef_[j]; }  // ENV:1886-1887
    }
    // ENV:1664-1682 (table mode)
    void gait_generator_table() {
        const float* row = ref + (size_t)ref_row() * 30;
        for (int j = 0; j < NJ; ++j) { jointRef_[j] = T(row[j]); jointDotRef_[j] = T(row[12 + j]); }
    }

    // ENV:1010-1109.  in_reset selects the RNG stream copy (reset() calls this twice, step once).
    void command_obs_update(bool flag_reset, uint32_t purpose) {
        if (flag_manual) return;
        if (flag_ManualTraj) {
            uint32_t r[4]; Philox::gen(seed, env_id, tick, purpose, r);
            float temp__ = u01(r[0]);
            if (temp__ < 0.5 / (max_time / control_dt_) || flag_reset) {            // ENV:1028
                temp__ = u01(r[1]);
                // temp__ < 0.2 : the "zero command" loop copies by value -> no-op (ENV:1039-1045, quirk 3)
                float u3 = u01(r[2]);
                if (0.2 < temp__ && temp__ <= 0.7) command[0] = T(u3) * Vx_max + (T(1.0) - T(u3)) * Vx_min;      // ENV:1046-1055
                else if (0.7 < temp__ && temp__ <= 0.85) command[1] = T(u3) * Vy_max + (T(1.0) - T(u3)) * Vy_min; // ENV:1058-1067
                else if (temp__ > 0.85) command[2] = T(u3) * omega_max + (T(1.0) - T(u3)) * omega_min;             // ENV:1068-1077
                // NB: temp__ == 0.2 exactly or temp__ < 0.2 falls in the reference's final else only when
                // !(0.2<t<=0.7) && !(0.7<t<=0.85): i.e. t<=0.2 ALSO reaches the omega branch (ENV:1056-1078).
                else command[2] = T(u3) * omega_max + (T(1.0) - T(u3)) * omega_min;
            }
            if (flag_reset) for (int i = 0; i < 3; ++i) command_filtered[i] = command[i];                           // ENV:1080-1085
            else for (int i = 0; i < 3; ++i) command_filtered[i] = command_filtered[i] * cmd_update_param + command[i] * (1 - cmd_update_param);  // ENV:1088-1092
            for (int i = 0; i < 3; ++i) obDouble_[i] = command_filtered[i];                                         // ENV:1095-1097
            gait_generator_manual(flag_reset);                                                                      // ENV:1098
        } else {
            const float* row = ref + (size_t)ref_row() * 30;                                                        // ENV:1102-1106
            for (int i = 0; i < 3; ++i) { obDouble_[i] = T(row[27 + i]); command_filtered[i] = obDouble_[i]; }
            gait_generator_table();
        }
    }

    // ENV:1116-1194
    void contact_obs_update() {
        if (!flag_TimeBasedContact) {
            for (int i = 0; i < 4; ++i) { contact_[i] = foot_in_contact[i] ? T(1) : T(0); contact_filtered[i] = contact_[i]; }
        } else {
            for (int i = 0; i < 4; ++i) {
                float real_phase = float(current_time() + phase_[i] * period_);             // ENV:1172-1181 (float locals)
                real_phase = float(mth::fmod(T(real_phase), period_) / period_);
                contact_filtered[i] = (T(real_phase) < lam_) ? T(1) : T(0); contact_[i] = contact_filtered[i];
            }
        }
    }

    // ENV:1199-1231
    void contact_information_update() {
        for (int i = 0; i < 4; ++i) {
            contact_force_norm[i] = 0;
            if (foot_in_contact[i]) {
                T n = mth::sqrt(foot_impulse[i][0] * foot_impulse[i][0] + foot_impulse[i][1] * foot_impulse[i][1] + foot_impulse[i][2] * foot_impulse[i][2]);
                contact_force_norm[i] = n / control_dt_;   // ENV:1208 (divides by control_dt, quirk 4)
            }
        }
        Kin k; kinematics(gc_, gv_, k);
        for (int i = 0; i < 4; ++i) contact_vel_norm[i] = mth::sqrt(dot(k.vtoe[i], k.vtoe[i]));   // ENV:1224-1231
    }

    // ENV:956-1004
    void updateObservation(uint32_t pbase) {
        for (int i = 0; i < 35; ++i) obDouble_[i] = 0;
        if (flag_manual || flag_ManualTraj) {
            obDouble_[3] = mth::sin(2 * T(BP5O_PI) * current_time() / period_);
            obDouble_[4] = mth::cos(2 * T(BP5O_PI) * current_time() / period_);
        } else {
            const float* row = ref + (size_t)ref_row() * 30; obDouble_[3] = T(row[25]); obDouble_[4] = T(row[26]);   // ENV:972
        }
        uint32_t r[4];
        for (int blk = 0; blk < 3; ++blk) {
            Philox::gen(seed, env_id, tick, pbase + P_OBS_Q0 + blk, r);
            for (int a = 0; a < 4; ++a) obDouble_[5 + 4 * blk + a] = (T(usym(r[a])) * jointNoise * noise_flag) + gc_[7 + 4 * blk + a];      // ENV:979
            Philox::gen(seed, env_id, tick, pbase + P_OBS_QD0 + blk, r);
            for (int a = 0; a < 4; ++a) obDouble_[17 + 4 * blk + a] = (T(usym(r[a])) * jointVelocityNoise * noise_flag) + gv_[6 + 4 * blk + a];  // ENV:983
        }
        M3<T> rot = quat_to_rot(gc_ + 3);   // ENV:986-993
        bodyFrameMatrix_ = rot;
        T g[4];
        gauss4(pbase + P_OBS_POSTURE, g);
        for (int a = 0; a < 3; ++a) obDouble_[29 + a] = rot.m[2][a] + (g[a] * noise_posture_sigma) * noise_flag;   // ENV:994-996
        V3<T> bl = tmul(rot, V3<T>(gv_[0], gv_[1], gv_[2])), ba = tmul(rot, V3<T>(gv_[3], gv_[4], gv_[5]));        // ENV:999-1000
        for (int a = 0; a < 3; ++a) { bodyLinearVel_[a] = bl[a]; bodyAngularVel_[a] = ba[a]; }
        gauss4(pbase + P_OBS_OMEGA, g);
        for (int a = 0; a < 3; ++a) obDouble_[32 + a] = ba[a] + noise_flag * (g[a] * noise_omega_sigma);           // ENV:1001-1003
    }
    // four standard normals by Box-Muller from one Philox block
    void gauss4(uint32_t purpose, T g[4]) const {
        uint32_t r[4]; Philox::gen(seed, env_id, tick, purpose, r);
        for (int p = 0; p < 2; ++p) {
            float u1 = (float)((r[2 * p] >> 8) + 1u) * (1.0f / 16777216.0f), u2 = u01(r[2 * p + 1]);
            T rad = mth::sqrt(T(-2.0) * mth::log(T(u1))), ang = T(6.283185307179586) * T(u2);
            g[2 * p] = rad * mth::cos(ang); g[2 * p + 1] = rad * mth::sin(ang);
        }
    }

    // ENV:1444-1548
    T DeepMimicRewardUpdate() {
        Kin k; kinematics(gc_, gv_, k);
        T ee = 0;
        for (int i = 0; i < 4; ++i) {
            V3<T> e = tmul(bodyFrameMatrix_, k.toe[i]);   // R^T (p_toe - p_base)  ENV:1452-1456
            for (int a = 0; a < 3; ++a) { EndEffector_[3 * i + a] = e[a]; T d = e[a] - EndEffectorRef_[3 * i + a]; ee += d * d; }
        }
        EndEffectorReward = EECoeff * mth::exp(-40 * ee);                                              // ENV:1459-1460
        T dz = gc_[2] - ground_height() - stand_height_;     // height above the terrain under the trunk (flat ground: z)
        BodyCenterReward = BodyPosCoeff * mth::exp(-80 * (dz * dz));                                   // ENV:1467-1476
        BodyAttitudeReward = BodyAttiCoeff * mth::exp(-80 * (obDouble_[29] * obDouble_[29] + obDouble_[30] * obDouble_[30]));  // ENV:1481-1483
        T jr = 0, jd = 0;
        for (int j = 0; j < NJ; ++j) { T a = jointRef_[j] - gc_[7 + j]; jr += a * a; T b = jointDotRef_[j] - gv_[6 + j]; jd += b * b; }
        JointReward = JointMimicCoeff * T(0.25) * mth::exp(T(-2.0) * jr);                              // ENV:1492-1493
        JointDotReward = JointMimicCoeff * T(0.75) * mth::exp(-control_dt_ * jd);                      // ENV:1494-1495
        T lref[3] = {flag_WildCat ? -command_filtered[0] : command_filtered[0], command_filtered[1], 0};   // ENV:1500-1502
        T aref[3] = {0, 0, command_filtered[2]};
        T le = 0, ae = 0; for (int a = 0; a < 3; ++a) { T d = bodyLinearVel_[a] - lref[a]; le += d * d; T e2 = bodyAngularVel_[a] - aref[a]; ae += e2 * e2; }
        VelocityReward = VelKeepCoeff / 2 * mth::exp(-2 * le) + VelKeepCoeff / 2 * mth::exp(-2 * ae);  // ENV:1504-1505
        T tn = 0, td = 0;
        for (int j = 0; j < NJ; ++j) { torque[j] = torque[j] / torque_limit[j]; tn += torque[j] * torque[j]; T d = torque[j] - torque_last[j]; td += d * d; }   // ENV:1511
        TorqueReward = TorqueCoeff / T(2.0) * mth::exp(T(-0.1) * tn) + TorqueCoeff / T(2.0) * mth::exp(T(-0.1) / control_dt_ * td);   // ENV:1513-1514
        for (int j = 0; j < NJ; ++j) torque_last[j] = torque[j];                                        // ENV:1515
        T cr = 0;
        for (int i = 0; i < 4; ++i) {                                                                   // ENV:1521-1528
            T real_phase = current_time() + phase_[i] * period_;
            real_phase = mth::fmod(real_phase, period_) / period_;
            cr += 4 * contact_vel_norm[i] * contact_vel_norm[i] * smooth_function(real_phase, T(2), lam_);
            cr += 2 * (contact_force_norm[i] / T(12.5)) * (contact_force_norm[i] / T(12.5)) * smooth_function2(real_phase, T(2), lam_);
        }
        ContactReward = ContactCoeff * mth::exp(-2 * cr);                                               // ENV:1529
        return (EndEffectorReward + BodyCenterReward + JointReward + JointDotReward + VelocityReward + BodyAttitudeReward + TorqueReward + ContactReward);  // ENV:1546-1547
    }

    // ENV:547-635
    void reset() {
        itera++;
        uint32_t r[4];
        Philox::gen(seed, env_id, tick, P_RST_TIME_CMD, r);
        t0_ = flag_manual ? T(0) : T(u01(r[3])); frame_idx = 0;                               // ENV:557, 565-573 (ManualTraj)
        if (!flag_ManualTraj && !flag_manual) {                                               // ENV:571
            double ratio = u01(r[2]);
            double rs = sampling_reshape(ratio);
            frame_idx = int((frame_max - frame_len - 10) * rs);
            t0_ = T(double(t0_) - double(frame_idx) * double(control_dt_));   // current_time_ (ENV:557,631,786) does not include the start row: current_time() adds frame_idx * dt back
        }
        for (int i = 0; i < 3; ++i) command_filtered[i] = 0;                                   // ENV:559-562
        for (int j = 0; j < NJ; ++j) torque_last[j] = 0;                                       // ENV:575
        command_obs_update(true, P_CMD + P_IN_RESET + 16);                                     // ENV:577
        for (int i = 0; i < 4; ++i) foot_in_contact[i] = 0;                                    // contact list is empty after setState
        contact_obs_update();                                                                  // ENV:579
        T random_init[NQ], random_vel_init[NV];
        for (int i = 0; i < NQ; ++i) random_init[i] = gc_init_[i];                             // ENV:327
        for (int i = 0; i < NV; ++i) random_vel_init[i] = 0;
        Philox::gen(seed, env_id, tick, P_RST_INIT, r);
        T n0 = T(usym(r[0])), n1 = T(usym(r[1]));
        // Eigen VectorXd * VectorXd evaluates a depth-1 product in Release: every joint gets the SAME
        // scalar noise[0]  ENV:583-586
        for (int j = 0; j < NJ; ++j) { random_init[7 + j] = jointRef_[j] * (n0 * T(0.3)) + jointRef_[j]; random_vel_init[6 + j] = jointDotRef_[j] * (n1 * T(0.3)) + jointDotRef_[j]; }
        Philox::gen(seed, env_id, tick, P_RST_BASEVEL, r);
        random_vel_init[0] = command_filtered[0] * (T(usym(r[0])) * T(0.2) + T(1.0));          // ENV:588
        random_vel_init[0] = flag_WildCat ? -random_vel_init[0] : random_vel_init[0];         // ENV:589
        random_vel_init[1] = command_filtered[1] * (T(usym(r[1])) * T(0.2) + T(1.0));          // ENV:590
        random_vel_init[5] = command_filtered[2] * (T(usym(r[2])) * T(0.2) + T(1.0));          // ENV:591
        if (!flag_manual) {                                                                    // ENV:593-606
            Philox::gen(seed, env_id, tick, P_RST_XY, r);
            float a = u01(r[0]), b = u01(r[1]);
            random_init[0] = T(a * 5.0f + (1.0f - a) * -5.0f); random_init[1] = T(b * 5.0f + (1.0f - b) * -5.0f);
        }
        if (flag_manual) { for (int i = 0; i < NQ; ++i) gc_[i] = gc_init_[i]; for (int i = 0; i < NV; ++i) gv_[i] = 0; }   // ENV:616-619
        else { for (int i = 0; i < NQ; ++i) gc_[i] = random_init[i]; for (int i = 0; i < NV; ++i) gv_[i] = random_vel_init[i]; }  // ENV:620-623
        if (terrain.valid()) gc_[2] += ground_height();   // drop height 0.35 m is measured from the terrain under the trunk (new spec)
        for (int i = 0; i < 4; ++i) { foot_impulse[i][0] = foot_impulse[i][1] = foot_impulse[i][2] = 0; }
        n_contacts = 0;
        updateObservation(P_IN_RESET);                                                         // ENV:625
        for (int i = 0; i < 35; ++i) obDouble_last_[i] = obDouble_[i];                         // ENV:626
        contact_obs_update();                                                                  // ENV:627
        command_obs_update(false, P_CMD + P_IN_RESET);                                         // ENV:628
        frame_idx++;                                                                           // ENV:630-631
        if (flag_crucial) meteor_respawn();                                                    // ENV:608-611
    }

    // ENV:827-838: spheres re-created above the robot (STATIC), size and mass grow with the episode time
    void meteor_respawn() {
        T t = current_time();
        met_r = (t / T(5.0) + T(1.0)) * T(0.08);                                               // cube_len ENV:1974
        met_m = T(num_cube) * (t / T(5) + T(0.2));
        met_p[0] = gc_[0] + T(0.05); met_p[1] = gc_[1]; met_p[2] = gc_[2] + T(1.0);             // circle_place(0, i, n, 1.0) ENV:836-837
        met_v[0] = met_v[1] = met_v[2] = 0; met_mode = 0;
    }
    // ENV:717-741 (schedule) + the sphere's own dynamics over one control step (new specification, see above)
    void meteor_update() {
        int every = int(T(5) * period_ / control_dt_);
        if (every > 0 && frame_idx % every == 0) { meteor_respawn(); return; }                  // ENV:731-735
        if (met_mode == 0) { met_mode = 1; met_v[0] = gv_[0]; met_v[1] = gv_[1]; met_v[2] = T(-5); }   // ENV:736-740, 849-858
        const T dt = simulation_dt_;
        int loopCount = int(control_dt_ / simulation_dt_ + T(1e-10));
        Kin k; kinematics(gc_, gv_, k);
        static thread_local T M[NV][NV]; Chol<T, NV> ch; bool have_m = false;
        T du[NV]; for (int i = 0; i < NV; ++i) du[i] = 0;
        const M3<T>& R = k.R[0];
        for (int sub = 0; sub < loopCount; ++sub) {
            met_v[2] -= model.gravity * dt;
            // ---- trunk box (centred on the trunk origin, URDF:26): closest point of the box to the sphere centre
            V3<T> dw(met_p[0] - gc_[0], met_p[1] - gc_[1], met_p[2] - gc_[2]);
            V3<T> d = transpose(R) * dw;
            V3<T> q(mth::min(mth::max(d.x, -model.box_half.x), model.box_half.x), mth::min(mth::max(d.y, -model.box_half.y), model.box_half.y),
                    mth::min(mth::max(d.z, -model.box_half.z), model.box_half.z));
            V3<T> del = d - q; T dist2 = dot(del, del);
            if (dist2 < met_r * met_r && dist2 > T(1e-12)) {
                V3<T> n = R * ((T(1) / mth::sqrt(dist2)) * del), x = R * q;
                V3<T> vb(gv_[0] + du[0], gv_[1] + du[1], gv_[2] + du[2]), wb(gv_[3] + du[3], gv_[4] + du[4], gv_[5] + du[5]);
                V3<T> vpt = vb + cross(wb, x);
                T vrel = n.x * (met_v[0] - vpt.x) + n.y * (met_v[1] - vpt.y) + n.z * (met_v[2] - vpt.z);
                if (vrel < T(0)) {
                    if (!have_m) { mass_matrix(k, M); if (!ch.factor(M)) throw std::runtime_error("mass matrix not PD"); have_m = true; }
                    T J[3][NV]; point_jacobian(k, 0, x, J);
                    T jn[NV], w[NV];
                    for (int c = 0; c < NV; ++c) jn[c] = J[0][c] * n.x + J[1][c] * n.y + J[2][c] * n.z;
                    ch.solve(jn, w);
                    T G = 0; for (int c = 0; c < NV; ++c) G += jn[c] * w[c];
                    T e = (-vrel > T(0.001)) ? T(0.95) : T(0);                                   // steel-steel pair ENV:244
                    T lam = -(T(1) + e) * vrel / (G + T(1) / met_m);
                    met_v[0] += lam / met_m * n.x; met_v[1] += lam / met_m * n.y; met_v[2] += lam / met_m * n.z;
                    for (int c = 0; c < NV; ++c) du[c] -= lam * w[c];
                }
            }
            // ---- ground: inelastic, Coulomb friction 0.8 on a point mass
            T hh = 0; V3<T> nn(0, 0, 1);
            if (terrain.valid()) terrain.sample(met_p[0], met_p[1], hh, nn);
            if ((met_p[2] - hh) * nn.z - met_r <= T(0)) {
                T vn = nn.x * met_v[0] + nn.y * met_v[1] + nn.z * met_v[2];
                if (vn < T(0)) {
                    met_v[0] -= vn * nn.x; met_v[1] -= vn * nn.y; met_v[2] -= vn * nn.z;
                    T vt = mth::sqrt(met_v[0] * met_v[0] + met_v[1] * met_v[1] + met_v[2] * met_v[2]);
                    if (vt > T(1e-9)) { T dv = mth::min(vt, T(0.8) * (-vn)) / vt; met_v[0] -= dv * met_v[0]; met_v[1] -= dv * met_v[1]; met_v[2] -= dv * met_v[2]; }
                }
            }
            for (int a = 0; a < 3; ++a) met_p[a] += met_v[a] * dt;
        }
        for (int i = 0; i < NV; ++i) gv_[i] += du[i];
    }

    // ENV:692-809
    T step(const float* action) {
        margin_geo = margin_rest = margin_term = margin_cone = T(1e30);
        uint32_t r[4]; Philox::gen(seed, env_id, tick, P_ACT, r);
        T an = T(usym(r[0]));
        for (int j = 0; j < NJ; ++j) {
            T p = T(action[j]); p = p * T(1.0) + actionMean_[j];                               // ENV:700-702
            p = (T(1.0) - filter_para) * p + filter_para * pTarget12Last_[j];                  // ENV:703
            p = p * (actionNoise * an) + p;                                                    // ENV:705 (shared scalar, quirk 16)
            pTarget12_[j] = p; pTarget12Last_[j] = p;                                          // ENV:706-707
        }
        int loopCount = int(control_dt_ / simulation_dt_ + T(1e-10));                          // ENV:711
        if (flag_crucial) meteor_update();                                                     // ENV:717-741
        // ENV:744-754: ForceDisturbance.  With Manual the base state is perturbed every 10 gait periods (state_disturbance,
        // ENV:912-940, ratio 0.5; the quaternion is re-normalised here, RaiSim's setState is assumed to do the same);
        // without Manual force_attack(random() < 0.0027) never fires: no external force.
        if (flag_ForceDisturbance && flag_manual) {
            int every = int(period_ / control_dt_ * T(10));
            if (every > 0 && frame_idx % every == 0) {
                uint32_t ra[4], rb[4]; Philox::gen(seed, env_id, tick, P_DISTURB, ra); Philox::gen(seed, env_id, tick, P_DISTURB + 1, rb);
                const T ratio = T(0.5);
                gc_[2] += T(0.03) * T(usym(ra[0])) * ratio;
                gc_[3] += T(0.1) * T(usym(ra[1])) * ratio; gc_[4] += T(0.1) * T(usym(ra[2])) * ratio;
                gc_[5] += T(0.1) * T(usym(ra[3])) * ratio; gc_[6] += T(0.1) * T(usym(rb[0])) * ratio;
                T qn = mth::sqrt(gc_[3] * gc_[3] + gc_[4] * gc_[4] + gc_[5] * gc_[5] + gc_[6] * gc_[6]);
                for (int i = 3; i < 7; ++i) gc_[i] /= qn;
                gv_[2] += T(0.1) * T(usym(rb[1])) * ratio; gv_[3] += T(0.3) * T(usym(rb[2])) * ratio; gv_[4] += T(0.3) * T(usym(rb[3])) * ratio;
            }
        }
        const T afp = T(0.99);                                                                 // ENV:756
        for (int i = 0; i < loopCount; ++i) {                                                  // ENV:758-774
            for (int j = 0; j < NJ; ++j) {
                torque[j] = (pTarget12_[j] - gc_[7 + j]) * jointPgain[j] - gv_[6 + j] * jointDgain[j];   // ENV:762-763
                torque[j] = afp * torque[j] + (T(1.0) - afp) * torque_last[j];                         // ENV:764
            }
            torque_clamp(gv_);                                                                 // ENV:765
            integrate(torque);                                                                 // ENV:766-768
        }
        updateObservation(0);                                                                  // ENV:776
        contact_information_update();                                                          // ENV:777
        T rew = DeepMimicRewardUpdate();                                                       // ENV:778
        command_obs_update(false, P_CMD);                                                      // ENV:784
        contact_obs_update();                                                                  // ENV:785
        frame_idx += 1;                                                                        // ENV:786-787
        return rew;
    }
    // ENV:1553-1578
    bool isTerminalState(float& terminalReward) const {
        terminalReward = float(terminalRewardCoeff_);
        T zrel = gc_[2] - ground_height();
        margin_term = mth::fmin(mth::fmin(mth::fabs(zrel - T(0.15)), mth::fabs(zrel - T(0.65))), mth::fabs(obDouble_[31] - T(0.5)));
        if (zrel < T(0.15) || zrel > T(0.65) || obDouble_[31] < T(0.5)) return true;
        terminalReward = 0.f; return false;
    }
    // ENV:1248-1262
    void observe(float* ob) {
        if (flag_ObsFilter) {
            for (int i = 5; i < 35; ++i) obDouble_[i] = obDouble_[i] * ObsFilterAlpha + obDouble_last_[i] * (T(1.0) - ObsFilterAlpha);
            for (int i = 0; i < 35; ++i) obDouble_last_[i] = obDouble_[i];
        }
        for (int i = 0; i < 35; ++i) ob[i] = float((obDouble_[i] - obMean_[i]) / obStd_[i]);
    }
    // ENV:942-950 in insertion order
    void extra_info(float* e) const {
        e[0] = float(EndEffectorReward); e[1] = float(BodyCenterReward); e[2] = float(gc_[2]);
        e[3] = float(BodyAttitudeReward); e[4] = float(JointReward); e[5] = float(VelocityReward);
    }
};

// ------------------------------------------------------------------ VectorizedEnvironment (VEC:127-382)
template <typename T> struct VecEnv {
    std::vector<Env<T>> envs;
    int num_threads = 1;
    uint32_t tick = 0;
    std::vector<float> ref; int ref_rows = 0;
    void create(const Cfg& c, int env_offset = 0, const double* compact_model = nullptr) {
        int n = (int)c.get("num_envs"); num_threads = (int)c.get("num_threads");
        uint32_t seed = (uint32_t)(int)c.get("seedd");   // VEC:171 (double truncated to int)
        envs.resize(n);
        if (compact_model) for (int i = 0; i < n; ++i) envs[i].model.set_compact(compact_model);
        for (int i = 0; i < n; ++i) envs[i].configure(c, (uint32_t)(env_offset + i), seed);
        if ((envs[0].flag_ManualTraj || envs[0].flag_manual) && !envs[0].flag_terrain) reset_all();   // VEC:172-182: init() resets every env once (tick 0); with a terrain the caller sets it first and then calls reset_all()
    }
    void set_ref(const float* data, int rows) {   // VEC:158-182
        ref.assign(data, data + (size_t)rows * 30); ref_rows = rows;
        for (auto& e : envs) { e.ref = ref.data(); e.ref_rows = rows; e.frame_max = rows / 2; e.frame_len = int(e.max_time / e.control_dt_); }   // ENV:538-539
        if (tick == 0 && !(envs[0].flag_ManualTraj || envs[0].flag_manual)) reset_all();   // table mode: the init reset needs the table
    }
    std::vector<float> hf;
    void set_terrain(const float* h, int nx, int ny, double x_size, double y_size, double cx, double cy) {
        hf.assign(h, h + (size_t)nx * ny);
        for (auto& e : envs) { e.terrain.h = hf.data(); e.terrain.nx = nx; e.terrain.ny = ny; e.terrain.dx = x_size / (nx - 1); e.terrain.dy = y_size / (ny - 1); e.terrain.cx = cx; e.terrain.cy = cy; }
    }
    void reset_all() {   // VEC:201-207 (serial in the reference)
        #pragma omp parallel for schedule(dynamic) num_threads(num_threads)
        for (int i = 0; i < (int)envs.size(); ++i) { envs[i].tick = tick; envs[i].reset(); }
        tick++;
    }
    void observe(float* ob) { for (size_t i = 0; i < envs.size(); ++i) envs[i].observe(ob + 35 * i); }
    // VEC:268-278 + perAgentStep VEC:352-372
    void step(const float* action, float* ob, float* reward, uint8_t* done, float* extra) {
        #pragma omp parallel for schedule(dynamic) num_threads(num_threads)
        for (int i = 0; i < (int)envs.size(); ++i) {
            Env<T>& e = envs[i]; e.tick = tick;
            reward[i] = float(e.step(action + 12 * i));
            float terminalReward = 0; done[i] = e.isTerminalState(terminalReward) ? 1 : 0;
            if (extra) e.extra_info(extra + 6 * i);
            if (done[i]) { e.reset(); reward[i] += terminalReward; }
            e.observe(ob + 35 * i);
        }
        tick++;
    }
};

}  // namespace bp5o
