// Minimal reader for the robot description the reference hands to RaiSim
// (ENV:231 world_->addArticulatedSystem(resourceDir_ + "/black_panther.urdf")): link inertials, joint origins / rotor inertia / damping,
// the trunk collision box and the toe collision sphere of the bp5 quadruped.  The step kernel works on a compact, mirror-symmetric
// model (one set of leg constants, signs per leg), so the four legs must be mirror images of each other; anything else is reported
// as an error instead of being silently approximated.  Pure host C++, no CUDA: irrl_parse_urdf() exposes it to CPU tests.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

namespace irrl {

struct UrdfLink { double mass = 0, com[3] = {0, 0, 0}, I[6] = {0, 0, 0, 0, 0, 0} /* ixx ixy ixz iyy iyz izz */; double box[3] = {0, 0, 0}, sphere_r = 0; bool found = false; };
struct UrdfJoint { double xyz[3] = {0, 0, 0}, axis[3] = {0, 0, 0}, damping = 0, rotor = 0; bool found = false; };

namespace urdf_detail {
inline std::string strip_comments(const std::string& s) {
    std::string o; size_t i = 0;
    while (i < s.size()) {
        size_t a = s.find("<!--", i);
        if (a == std::string::npos) { o.append(s, i, std::string::npos); break; }
        o.append(s, i, a - i);
        size_t b = s.find("-->", a + 4);
        if (b == std::string::npos) break;
        i = b + 3;
    }
    return o;
}
// value of attribute `attr` inside the first `<tag ...>` found in `block` ("" if absent)
inline std::string attr_of(const std::string& block, const std::string& tag, const std::string& attr) {
    size_t p = 0;
    while ((p = block.find("<" + tag, p)) != std::string::npos) {
        const char c = p + 1 + tag.size() < block.size() ? block[p + 1 + tag.size()] : ' ';
        if (c != ' ' && c != '\t' && c != '\n' && c != '/' && c != '>') { p += 1; continue; }      // e.g. <origin vs <originx
        size_t e = block.find('>', p); if (e == std::string::npos) return "";
        const std::string el = block.substr(p, e - p);
        size_t a = el.find(attr + "=\"");
        while (a != std::string::npos && a > 0 && el[a - 1] != ' ' && el[a - 1] != '\t' && el[a - 1] != '\n') a = el.find(attr + "=\"", a + 1);
        if (a == std::string::npos) return "";
        a += attr.size() + 2; size_t q = el.find('"', a);
        return q == std::string::npos ? "" : el.substr(a, q - a);
    }
    return "";
}
inline int numbers(const std::string& s, double* out, int n) { std::istringstream is(s); int k = 0; while (k < n && (is >> out[k])) ++k; return k; }
// body of <kind name="name" ...> ... </kind>
inline bool element(const std::string& doc, const std::string& kind, const std::string& name, std::string& head, std::string& body) {
    size_t p = 0; const std::string key = "name=\"" + name + "\"";
    while ((p = doc.find("<" + kind + " ", p)) != std::string::npos) {
        size_t e = doc.find('>', p); if (e == std::string::npos) return false;
        head = doc.substr(p, e - p + 1);
        if (head.find(key) != std::string::npos) {
            if (head.size() >= 2 && head[head.size() - 2] == '/') { body.clear(); return true; }
            size_t c = doc.find("</" + kind + ">", e); if (c == std::string::npos) return false;
            body = doc.substr(e + 1, c - e - 1); return true;
        }
        p = e;
    }
    return false;
}
inline std::string section(const std::string& body, const std::string& tag) {
    size_t a = body.find("<" + tag); if (a == std::string::npos) return "";
    size_t b = body.find("</" + tag + ">", a); return b == std::string::npos ? "" : body.substr(a, b - a);
}
}  // namespace urdf_detail

inline bool urdf_link(const std::string& doc, const std::string& name, UrdfLink& L) {
    using namespace urdf_detail;
    std::string head, body; if (!element(doc, "link", name, head, body)) return false;
    L.found = true;
    const std::string in = section(body, "inertial");
    if (!in.empty()) {
        L.mass = atof(attr_of(in, "mass", "value").c_str());
        numbers(attr_of(in, "origin", "xyz"), L.com, 3);
        const char* k[6] = {"ixx", "ixy", "ixz", "iyy", "iyz", "izz"};
        for (int i = 0; i < 6; ++i) L.I[i] = atof(attr_of(in, "inertia", k[i]).c_str());
    }
    const std::string col = section(body, "collision");
    if (!col.empty()) { numbers(attr_of(col, "box", "size"), L.box, 3); L.sphere_r = atof(attr_of(col, "sphere", "radius").c_str()); }
    return true;
}
inline bool urdf_joint(const std::string& doc, const std::string& name, UrdfJoint& J) {
    using namespace urdf_detail;
    std::string head, body; if (!element(doc, "joint", name, head, body)) return false;
    J.found = true;
    numbers(attr_of(body, "origin", "xyz"), J.xyz, 3); numbers(attr_of(body, "axis", "xyz"), J.axis, 3);
    J.damping = atof(attr_of(body, "dynamics", "damping").c_str()); J.rotor = atof(attr_of(body, "dynamics", "rotor_inertia").c_str());
    return true;
}

// The compact model of the kernels (EnvParams fields, irrl_params.h), in the canonical front-left frame (sx = sy = +1).
struct CompactModel {
    double I0[3], I1[3], I2[4], I3[3], rotor[3], off1x, off1y, off2y, toe_z, toe_r, box_half[3], joint_damping;
    double m0, com0[3], m1, com1[3], m2, com2[3], m3, com3z, knee_z;
};

// Reads `path`; returns 0 and fills M, or a negative code with a message in err.
inline int read_urdf(const std::string& path, CompactModel& M, std::string& err) {
    std::ifstream f(path); if (!f.is_open()) { err = "cannot open " + path; return -1; }
    std::stringstream ss; ss << f.rdbuf();
    const std::string doc = urdf_detail::strip_comments(ss.str());
    UrdfLink body; if (!urdf_link(doc, "body", body)) { err = "link 'body' missing"; return -2; }
    const char* legs[4] = {"fr", "fl", "hr", "hl"};
    UrdfLink ab[4], th[4], sh[4], toe[4]; UrdfJoint j1[4], j2[4], j3[4], jt[4];
    for (int l = 0; l < 4; ++l) {
        const std::string s = legs[l];
        if (!urdf_link(doc, "abduct_" + s, ab[l]) || !urdf_link(doc, "thigh_" + s, th[l]) || !urdf_link(doc, "shank_" + s, sh[l]) || !urdf_link(doc, "toe_" + s, toe[l]) ||
            !urdf_joint(doc, "torso_to_abduct_" + s + "_j", j1[l]) || !urdf_joint(doc, "abduct_" + s + "_to_thigh_" + s + "_j", j2[l]) ||
            !urdf_joint(doc, "thigh_" + s + "_to_knee_" + s + "_j", j3[l]) || !urdf_joint(doc, "toe_" + s + "_joint", jt[l])) {
            err = "leg '" + s + "': a link or joint of the bp5 naming scheme is missing"; return -2;
        }
    }
    // mirror symmetry against the front-left leg: x flips front/hind, y flips left/right (inertia products flip with the axis they contain)
    auto close = [](double a, double b) { return std::fabs(a - b) <= 1e-9 + 1e-6 * std::fabs(b); };
    const int FL = 1;
    for (int l = 0; l < 4; ++l) {
        const double sx = (l < 2) ? 1.0 : -1.0, sy = (l & 1) ? 1.0 : -1.0;
        const UrdfLink* A[3] = {&ab[l], &th[l], &sh[l]}; const UrdfLink* B[3] = {&ab[FL], &th[FL], &sh[FL]};
        bool ok = true;
        for (int k = 0; k < 3; ++k) {
            ok = ok && close(A[k]->mass, B[k]->mass) && close(A[k]->com[0], sx * B[k]->com[0]) && close(A[k]->com[1], sy * B[k]->com[1]) && close(A[k]->com[2], B[k]->com[2]);
            ok = ok && close(A[k]->I[0], B[k]->I[0]) && close(A[k]->I[3], B[k]->I[3]) && close(A[k]->I[5], B[k]->I[5]);
            ok = ok && close(A[k]->I[1], sx * sy * B[k]->I[1]) && close(A[k]->I[2], sx * B[k]->I[2]) && close(A[k]->I[4], sy * B[k]->I[4]);
        }
        ok = ok && close(toe[l].mass, toe[FL].mass) && close(toe[l].sphere_r, toe[FL].sphere_r) && close(toe[l].I[0], toe[FL].I[0]);
        ok = ok && close(j1[l].xyz[0], sx * j1[FL].xyz[0]) && close(j1[l].xyz[1], sy * j1[FL].xyz[1]) && close(j1[l].xyz[2], j1[FL].xyz[2]);
        ok = ok && close(j2[l].xyz[0], sx * j2[FL].xyz[0]) && close(j2[l].xyz[1], sy * j2[FL].xyz[1]) && close(j2[l].xyz[2], j2[FL].xyz[2]);
        ok = ok && close(j3[l].xyz[2], j3[FL].xyz[2]) && close(jt[l].xyz[2], jt[FL].xyz[2]);
        ok = ok && close(j1[l].rotor, j1[FL].rotor) && close(j2[l].rotor, j2[FL].rotor) && close(j3[l].rotor, j3[FL].rotor) && close(j1[l].damping, j1[FL].damping);
        if (!ok) { err = std::string("leg '") + legs[l] + "' is not the mirror image of leg 'fl': the compact model of the kernels cannot represent it"; return -3; }
    }
    // structure the kernels assume: abad about x, hip / knee about -y, joint offsets (x, y, 0), (0, y, 0), (0, 0, z), toe on the shank axis
    const UrdfJoint &a = j1[FL], &h = j2[FL], &k = j3[FL], &t = jt[FL];
    if (!(close(a.axis[0], 1) && close(a.axis[1], 0) && close(a.axis[2], 0) && close(h.axis[1], -1) && close(h.axis[0], 0) && close(k.axis[1], -1) && close(k.axis[0], 0)) ||
        !(close(a.xyz[2], 0) && close(h.xyz[0], 0) && close(h.xyz[2], 0) && close(k.xyz[0], 0) && close(k.xyz[1], 0) && close(t.xyz[0], 0) && close(t.xyz[1], 0)) ||
        !(close(body.I[1], 0) && close(body.I[2], 0) && close(body.I[4], 0) && close(ab[FL].I[1], 0) && close(ab[FL].I[2], 0) && close(ab[FL].I[4], 0) &&
          close(th[FL].I[1], 0) && close(th[FL].I[2], 0) && close(sh[FL].I[1], 0) && close(sh[FL].I[2], 0) && close(sh[FL].I[4], 0)) ||
        !(close(sh[FL].com[0], 0) && close(sh[FL].com[1], 0) && close(toe[FL].com[0], 0) && close(toe[FL].com[1], 0) && close(toe[FL].com[2], 0))) {
        err = "joint axes / offsets / inertia products outside the bp5 structure the kernels implement"; return -4;
    }
    for (int i = 0; i < 3; ++i) { M.com0[i] = body.com[i]; M.box_half[i] = 0.5 * body.box[i]; M.com1[i] = ab[FL].com[i]; M.com2[i] = th[FL].com[i]; }
    M.m0 = body.mass; M.I0[0] = body.I[0]; M.I0[1] = body.I[3]; M.I0[2] = body.I[5];
    M.m1 = ab[FL].mass; M.I1[0] = ab[FL].I[0]; M.I1[1] = ab[FL].I[3]; M.I1[2] = ab[FL].I[5];
    M.m2 = th[FL].mass; M.I2[0] = th[FL].I[0]; M.I2[1] = th[FL].I[3]; M.I2[2] = th[FL].I[5]; M.I2[3] = -th[FL].I[4];      // stored as |iyz| of the right legs
    // shank and toe are one rigid body (fixed joint): merged about the common centre of mass
    const double ms = sh[FL].mass, mt = toe[FL].mass, zs = sh[FL].com[2], zt = t.xyz[2];
    const double zc = (ms * zs + mt * zt) / (ms + mt), ds = zs - zc, dt = zt - zc;
    M.m3 = ms + mt; M.com3z = zc;
    M.I3[0] = sh[FL].I[0] + ms * ds * ds + toe[FL].I[0] + mt * dt * dt;
    M.I3[1] = sh[FL].I[3] + ms * ds * ds + toe[FL].I[3] + mt * dt * dt;
    M.I3[2] = sh[FL].I[5] + toe[FL].I[5];
    M.rotor[0] = a.rotor; M.rotor[1] = h.rotor; M.rotor[2] = k.rotor;
    M.off1x = a.xyz[0]; M.off1y = a.xyz[1]; M.off2y = h.xyz[1]; M.knee_z = k.xyz[2]; M.toe_z = t.xyz[2]; M.toe_r = toe[FL].sphere_r;
    M.joint_damping = a.damping;
    if (!(M.m0 > 0 && M.m1 > 0 && M.m2 > 0 && M.m3 > 0 && M.toe_r > 0 && M.box_half[0] > 0)) { err = "non-positive mass / collision size"; return -5; }
    return 0;
}

}  // namespace irrl
