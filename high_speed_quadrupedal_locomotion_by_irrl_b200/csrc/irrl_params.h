// Plain-old-data parameter block and device-state layout shared by the host shim and the sm_100a kernels.
// Citations: ENV = flex_gym/env/env/BlackPanther_V55/Environment.hpp, VEC = flex_gym/env/VectorizedEnvironment.hpp,
// URDF = .../BlackPanther_V55/urdf/black_panther.urdf (all under /root/reference/IRRL/FlexibleRobotRaisimGym).
#pragma once
#include <cstdint>

namespace irrl {

constexpr int OB_DIM = 35;       // ENV:360
constexpr int ACT_DIM = 12;      // ENV:361
constexpr int EXTRA_DIM = 6;     // ENV:944-949
constexpr int GC_DIM = 19;       // ENV:295
constexpr int GV_DIM = 18;       // ENV:296
constexpr int STATE_DIM = 192;   // flat per-env state vector of irrl_get_state / irrl_set_state (include/irrl_b200.h)
constexpr int LSTM_H = 48;       // run_bp_v5.py:111
constexpr int LSTM_STATE = 384;  // run_bp_v5.py:136-137

// RNG stream ids (Philox4x32-10 counter = (global env id, tick, purpose, 0), key = (seed, 0x1BD11BDA))
enum : uint32_t {
    P_OBS_Q0 = 0, P_OBS_QD0 = 3, P_OBS_POSTURE = 6, P_OBS_OMEGA = 7, P_CMD = 8, P_ACT = 9,
    P_RST_TIME_CMD = 16, P_RST_INIT = 17, P_RST_BASEVEL = 18, P_RST_XY = 19, P_DISTURB = 20, P_IN_RESET = 32,
    P_DR_MATERIAL = 64, P_DR_MASS = 65, P_DR_COM = 80, P_DR_CALF = 96,
    P_POLICY_EPS = 128,   // +0..2 : 12 standard normals of the Gaussian policy sample
};

struct EnvParams {
    // ---- YAML configuration (ENV:1594-1659, VEC:146-171)
    float abad, period, lam, stand_height, up_height_max;
    float Vx_max, Vx_min, Vy_max, Vy_min, omega_max, omega_min, lean_front, lean_hind;
    int flag_manual, flag_filter, flag_height_variable, flag_time_contact, flag_manual_traj, flag_obs_filter,
        flag_wildcat, flag_force_dist, flag_terrain, flag_stochastic;
    float terminal_coeff, ee_coeff, pos_coeff, atti_coeff, joint_coeff, vel_coeff, torque_coeff, contact_coeff;
    float stiffness, abad_ratio, damping, max_time, action_noise, noise_flag;
    float motor_max_torque, motor_crit_speed, motor_max_speed;
    float sim_dt, control_dt;
    double control_dt_d, period_d, inv_period_d; // the YAML doubles: the gait phase of the contact reward / time-based contact flags is reduced in double (a handful of fp64 operations per control step)
    int loop_count;               // ENV:711
    int disturb_every;            // ENV:746 : control steps between state disturbances = int(period / control_dt * 10)
    int flag_crucial, num_cube, meteor_every;   // Crutial: True (ENV:717-741, 815-861): meteor sphere re-created every int(5 * period / control_dt) control steps
    int gait_type;
    float phase[4];               // ENV:398-409
    float filter_para;            // ENV:396
    float obs_filter_alpha;       // ENV:425-426
    // ---- hard-coded members of the reference (ENV:1929-2083)
    float l_thigh, l_calf, l_hip, max_len;
    float joint_noise, joint_vel_noise, posture_sigma, omega_sigma, cmd_update;
    // ---- robot model (URDF)
    float I0[3];                  // trunk inertia diag              URDF:21
    float I1[3];                  // abad link inertia diag          URDF:64
    float I2[4];                  // thigh xx,yy,zz,|yz|             URDF:92
    float I3[3];                  // shank+toe merged about its COM  URDF:118,155
    float rotor[3];               // URDF:56,84,110
    float off1x, off1y, off2y;    // |joint offsets|                 URDF:52,80
    float toe_z, toe_r;           // URDF:162,148
    float box_half[3];            // URDF:26
    float gravity, joint_damping;
    // nominal per-body mass/com and knee offset (per-env copies live in DevState::legmodel / basemodel)
    float m0, com0[3], m1, com1[3], m2, com2[3], m3, com3z, knee_z;
    float mu, restitution, rest_threshold;   // default material ENV:433
    // ---- contact solver (new specification)
    int solver_iters, slide_iters, jacobi_sweeps;
    float solver_tol;
    // ---- bookkeeping
    uint32_t seed, env_offset;
    int N;
    const float* ref;             // optional [rows,30] reference table (ENV:17-21), device pointer
    int ref_rows, frame_max, frame_len;
    // heightfield terrain (Terrain: True, ENV:252-265): [nx][ny] heights, grid centred on (cx, cy), device pointer
    const float* terrain; int terrain_nx, terrain_ny; float terrain_cx, terrain_cy, terrain_dx, terrain_dy;
};

// Structure-of-arrays persistent state in HBM.  One robot is served by a group of 4 lanes (one per leg);
// every per-leg block is 16-byte aligned so a lane moves its leg with float4 transactions.
struct DevState {
    float* base;       // [N][16] : p(3) quat wxyz(4) v(3) w(3) t0 pad(2)
    float* cmd;        // [N][8]  : command(3) command_filtered(3) pad(2)
    float* legs;       // [N][4][16] : q(3) qd(3) torque_last(3) jointRef(3) jointDotRef(3) pad
    float* legs2;      // [N][4][12] : pTargetLast(3) EndEffectorRef(3) applied torque(3) contact flag, |impulse|, pad
    float* obd;        // [N][36] : unscaled observation obDouble_ (ENV:1932)
    float* obd_last;   // [N][36] : obDouble_last_ (ObsFilter only)
    float* legmodel;   // [N][4][16] : m1 m2 m3 knee_z | com1 pad | com2 pad | com3 pad
    float* basemodel;  // [N][8]  : m0 com0(3) mu restitution threshold pad
    int* frame_idx;    // [N]
    int* itera;        // [N] reset counter (ENV:554)
    int* ep_len;       // [N] steps since last reset (RaisimGymVecEnv.py:42-50 bookkeeping)
    float* ep_ret;     // [N] reward sum since last reset
    int* solver_sweeps;// [N] Gauss-Seidel sweeps used by the last substep (diagnostic)
    float* meteor;     // [N][12] : sphere p(3) v(3) mode radius mass pad(3)  (Crutial: True only, else null)
};

}  // namespace irrl
