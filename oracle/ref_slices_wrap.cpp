// TEST INFRASTRUCTURE.  C entry points over source slices of the REFERENCE's own Environment.hpp, cut out at build time by
// oracle/Makefile (target `ref`) from /root/reference into oracle/_ref/*.inc (git-ignored: reference sources never enter
// this repository) and compiled into oracle/_ref/libenv_ref_slices.so.  The reference's environment as a whole cannot be built
// here (RaiSim, Eigen, yaml-cpp, OGRE are absent), but these pieces are plain double arithmetic:
//   free functions   ENV:45 (PI), 71-81 sampling_reshape, 96-99 gauss, 118-156 smooth_function / smooth_function2, 169-179 sgn
//   member function  ENV:1687-1751 inverse_kinematics (needs only the member max_len)
//   statement block  ENV:1275-1279 + 1289-1305, the arithmetic of torque_clamp (vectors replaced by plain arrays of the same names)
// tests/test_oracle_ref_slices.py checks the oracle's restatements against them and against the golden vectors they produced
// (tests/golden/ref_slices.npz, written by oracle/gen_ref_slices_golden.py).
#include <cmath>
#include <cstdlib>
#include <iostream>
using namespace std;

#include "env_free_functions.inc"

namespace {
struct RefEnvSlice {
    double max_len;
#include "env_inverse_kinematics.inc"
};
}  // namespace

extern "C" {
double ref_sampling_reshape(double ratio) { return sampling_reshape(ratio); }
double ref_gauss(double x, double width, double height) { return gauss(x, width, height); }
double ref_smooth_function(double phase, double slope, double lam) { return smooth_function(phase, slope, lam); }
double ref_smooth_function2(double phase, double slope, double lam) { return smooth_function2(phase, slope, lam); }
double ref_sgn(double x) { return sgn(x); }
double ref_pi() { return PI; }
void ref_inverse_kinematics(double x, double y, double z, double l_hip, double l_thigh, double l_calf, double max_len, double* theta, int is_right) {
    RefEnvSlice e; e.max_len = max_len;
    e.inverse_kinematics(x, y, z, l_hip, l_thigh, l_calf, theta, is_right != 0);
}
// torque[12] in/out, gv_temp[18], upper/lower[18] out (ENV:1273-1305)
void ref_torque_clamp(double* torque, const double* gv_temp, double MotorMaxTorque, double MotorCriticalSpeed, double MotorMaxSpeed, double* upper, double* lower) {
    const int nJoints_ = 12;
#include "env_torque_clamp_head.inc"
#include "env_torque_clamp_loop.inc"
}
}
