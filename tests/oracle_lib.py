"""ctypes access to the CPU oracle (oracle/libbp5_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libbp5_oracle.so")
_lib = None
STATE_DIM = 192
# slices of the flat per-env state vector shared by the oracle and the CUDA library (include/irrl_b200.h)
S = dict(gc=slice(0, 19), gv=slice(19, 37), ptarget_last=slice(37, 49), torque_last=slice(49, 61),
         command=slice(61, 64), command_filtered=slice(64, 67), joint_ref=slice(67, 79), joint_dot_ref=slice(79, 91),
         ee_ref=slice(91, 103), t0=103, frame_idx=104, contact=slice(105, 109), ob=slice(109, 144),
         ob_last=slice(144, 179), torque=slice(179, 191), itera=191)


def build():
    src = [os.path.join(ROOT, "oracle", f) for f in ("oracle_capi.cpp", "bp5_oracle.hpp")]
    if os.path.exists(_SO) and os.path.exists(_SO_FAST) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src if os.path.exists(s)):
        return _SO
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


_SO_FAST = os.path.join(ROOT, "oracle", "libbp5_oracle_fast.so")   # same source, the reference's build flags (-O3 -mtune=native, CMakeLists.txt:28): timed CPU baseline only
_lib_fast = None


def lib(fast=False):
    global _lib, _lib_fast
    if fast:
        if _lib_fast is None:
            build()
            _lib_fast = _declare(C.CDLL(_SO_FAST))
        return _lib_fast
    if _lib is None:
        build()
        _lib = _declare(C.CDLL(_SO))
    return _lib


def _declare(L):
    if True:
        L.bp5o_create.restype = C.c_void_p
        L.bp5o_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.bp5o_time_steps.restype = C.c_double
        L.bp5o_time_steps.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint, C.POINTER(C.c_long)]
        L.bp5o_gait.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p, C.c_int] + [C.c_void_p] * 3
        for name in ("destroy", "reset", "observe", "step", "get_state", "set_state", "mass_and_h", "body_kin", "toe_kin",
                     "integrate", "contact_info", "reward_terms", "margins", "model_params", "set_ref", "set_tick"):
            getattr(L, "bp5o_" + name).argtypes = None
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def count_flops(cfg: dict, warm=200, steps=300, sigma=0.1, seed=0):
    """Floating-point operations per env-step of the oracle itself (Counted scalar): (total, dict of op classes)"""
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import to_kv_string
    L = lib()
    L.bp5o_count_flops.restype = C.c_double
    L.bp5o_count_flops.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_void_p]
    out = np.zeros(6)
    total = L.bp5o_count_flops(to_kv_string(dict(cfg, num_threads=1)).encode(), warm, steps, sigma, seed, out.ctypes.data)
    return total, dict(zip(("add", "mul", "div", "sqrt", "transcendental", "compare"), out))


class Oracle:
    """The reference-semantics CPU vec-env: precision 'double' (like the reference) or 'float'."""

    def __init__(self, cfg: dict, precision="double", env_offset=0, model40=None, fast=False):
        from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import to_kv_string
        self.L = lib(fast)
        if model40 is None:
            self.h = C.c_void_p(self.L.bp5o_create(to_kv_string(cfg).encode(), 0 if precision == "double" else 1, env_offset))
        else:   # a robot description other than the shipped one: the 40 numbers of irrl_parse_urdf (include/irrl_b200.h)
            m = np.ascontiguousarray(model40, np.float64); assert m.shape == (40,)
            self.L.bp5o_create_with_model.restype = C.c_void_p
            self.h = C.c_void_p(self.L.bp5o_create_with_model(to_kv_string(cfg).encode(), 0 if precision == "double" else 1, env_offset, _p(m)))
        if not self.h:
            raise RuntimeError("oracle creation failed (missing cfg key?)")
        self.n = self.L.bp5o_num_envs(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.bp5o_destroy(self.h)
        except Exception:
            pass

    def reset(self):
        ob = np.zeros((self.n, 35), np.float32)
        self.L.bp5o_reset(self.h, _p(ob))
        return ob

    def observe(self):
        ob = np.zeros((self.n, 35), np.float32)
        self.L.bp5o_observe(self.h, _p(ob))
        return ob

    def step(self, action):
        action = np.ascontiguousarray(action, np.float32)
        ob = np.zeros((self.n, 35), np.float32)
        rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.uint8)
        extra = np.zeros((self.n, 6), np.float32)
        rc = self.L.bp5o_step(self.h, _p(action), _p(ob), _p(rew), _p(done), _p(extra))
        if rc != 0:
            raise RuntimeError("oracle step failed")
        return ob, rew, done.astype(bool), extra

    def set_tick(self, t):
        self.L.bp5o_set_tick(self.h, C.c_uint(t))

    def get_tick(self):
        return self.L.bp5o_get_tick(self.h)

    def get_state(self, env=None):
        if env is None:
            return np.stack([self.get_state(i) for i in range(self.n)])
        s = np.zeros(STATE_DIM, np.float64)
        self.L.bp5o_get_state(self.h, C.c_int(env), _p(s))
        return s

    def set_state(self, env, s):
        s = np.ascontiguousarray(s, np.float64)
        assert s.shape == (STATE_DIM,)
        self.L.bp5o_set_state(self.h, C.c_int(env), _p(s))

    def get_meteor(self):
        """[n,9]: sphere position, velocity, mode (0 static / 1 falling), radius, mass (Crutial: True)"""
        out = np.zeros((self.n, 9), np.float64)
        for i in range(self.n):
            self.L.bp5o_get_meteor(self.h, C.c_int(i), _p(out[i]))
        return out

    def set_meteor(self, m):
        m = np.ascontiguousarray(m, np.float64)
        for i in range(self.n):
            self.L.bp5o_set_meteor(self.h, C.c_int(i), _p(m[i]))

    def mass_and_h(self, env=0):
        M = np.zeros((18, 18)); h = np.zeros(18)
        self.L.bp5o_mass_and_h(self.h, C.c_int(env), _p(M), _p(h))
        return M, h

    def body_kin(self, env=0):
        out = np.zeros((13, 28))
        self.L.bp5o_body_kin(self.h, C.c_int(env), _p(out))
        return dict(R=out[:, 0:9].reshape(13, 3, 3), pc=out[:, 9:12], vc=out[:, 12:15], w=out[:, 15:18], mass=out[:, 18],
                    I=out[:, 19:28].reshape(13, 3, 3))

    def toe_kin(self, env=0):
        out = np.zeros((4, 6))
        self.L.bp5o_toe_kin(self.h, C.c_int(env), _p(out))
        return out[:, :3], out[:, 3:]

    def gait(self, t, cmd, is_first=False, env=0):
        cmd = np.ascontiguousarray(cmd, np.float64)
        a, b, c = np.zeros(12), np.zeros(12), np.zeros(12)
        self.L.bp5o_gait(self.h, env, float(t), _p(cmd), int(is_first), _p(a), _p(b), _p(c))
        return a, b, c

    def integrate(self, env, tau12):
        tau = np.ascontiguousarray(tau12, np.float64)
        if self.L.bp5o_integrate(self.h, C.c_int(env), _p(tau)) != 0:
            raise RuntimeError("oracle integrate failed")

    def contact_info(self, env=0):
        o = np.zeros(26)
        self.L.bp5o_contact_info(self.h, C.c_int(env), _p(o))
        return dict(foot_in_contact=o[0:4].astype(int), foot_impulse=o[4:16].reshape(4, 3), n_contacts=int(o[16]),
                    sweeps=int(o[17]), force_norm=o[18:22], vel_norm=o[22:26])

    def reward_terms(self, env=0):
        o = np.zeros(8)
        self.L.bp5o_reward_terms(self.h, C.c_int(env), _p(o))
        return o

    def margins(self):
        """[n,4]: how far the last control step's discrete decisions were from their thresholds (geometric gap [m], restitution
        threshold [m/s], termination bounds, stick/slide boundary [relative]).  Captured by step() BEFORE the auto-reset."""
        out = np.zeros((self.n, 4))
        for i in range(self.n):
            self.L.bp5o_margins(self.h, C.c_int(i), _p(out[i]))
        return out

    def model_params(self, env=0):
        o = np.zeros(3 + 13 * 7)
        self.L.bp5o_model_params(self.h, C.c_int(env), _p(o))
        return dict(mu=o[0], restitution=o[1], threshold=o[2], mass=o[3:].reshape(13, 7)[:, 0], com=o[3:].reshape(13, 7)[:, 1:4],
                    off=o[3:].reshape(13, 7)[:, 4:7])

    def is_terminal(self, env=0):
        return bool(self.L.bp5o_is_terminal(self.h, C.c_int(env)))

    def time_steps(self, steps, sigma=0.0, seed=0):
        nd = C.c_long(0)
        t = self.L.bp5o_time_steps(self.h, int(steps), float(sigma), int(seed), C.byref(nd))
        return t, nd.value
