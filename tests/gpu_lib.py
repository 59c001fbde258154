"""Helpers for the -m gpu parity tests: the CUDA env through its C ABI (host numpy buffers) next to the oracle."""
import numpy as np

from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import dump_yaml


class Cuda:
    def __init__(self, cfg: dict, env_offset=0, device=0):
        self.env = FlexibleGymEnv("", dump_yaml(cfg), device=device, env_offset=env_offset)
        self.env.init()
        self.n = self.env.getNumOfEnvs()

    def reset(self):
        ob = np.zeros((self.n, 35), np.float32)
        self.env.reset(ob)
        return ob

    def observe(self):
        ob = np.zeros((self.n, 35), np.float32)
        self.env.observe(ob)
        return ob

    def step(self, action):
        action = np.ascontiguousarray(action, np.float32)
        ob = np.zeros((self.n, 35), np.float32); rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.bool_); extra = np.zeros((self.n, 6), np.float32)
        self.env.step(action, ob, rew, done, extra)
        return ob, rew, done, extra

    def get_state(self):
        s = np.zeros((self.n, 192), np.float32)
        self.env.getState(s)
        return s

    def set_state(self, s):
        self.env.setState(np.ascontiguousarray(s, np.float32))

    def mass_matrix(self):
        m = np.zeros((self.n, 324), np.float32); self.env.GetMassMatrix(m); return m.reshape(self.n, 18, 18)

    def inverse_mass_matrix(self):
        m = np.zeros((self.n, 324), np.float32); self.env.GetInverseMassMatrix(m); return m.reshape(self.n, 18, 18)

    def nonlinear(self):
        h = np.zeros((self.n, 18), np.float32); self.env.GetNonlinear(h); return h

    def integrate(self, tau):
        c = np.zeros((self.n, 16), np.float32)
        self.env.integrate(np.ascontiguousarray(tau, np.float32), c)
        c = c.reshape(self.n, 4, 4)
        return c[:, :, 0].astype(int), c[:, :, 1:]

    def sweeps(self):
        s = np.zeros(self.n, np.int32); self.env.getSolverSweeps(s); return s

    def get_meteor(self):
        m = np.zeros((self.n, 9), np.float32); self.env.getMeteor(m); return m

    def set_meteor(self, m):
        self.env.setMeteor(np.ascontiguousarray(m, np.float32))

    def model_params(self):
        o = np.zeros((self.n, 94), np.float32); self.env.getModelParams(o); return o


def rel(a, b):
    """norm-wise relative error max|a-b| / max|b| (per call), the measure used for the 1e-5 / 1e-4 bars"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def random_states(rng, n, z=(0.5, 0.6), vel=1.0, jitter=0.5):
    from oracle_lib import S, STATE_DIM
    s = np.zeros((n, STATE_DIM))
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    s[:, 0:2] = rng.uniform(-5, 5, size=(n, 2)); s[:, 2] = rng.uniform(z[0], z[1], size=n)
    s[:, 3:7] = q
    s[:, 7:19] = np.tile([0.0, -0.78, 1.57], 4) + rng.uniform(-jitter, jitter, size=(n, 12))
    s[:, 19:22] = rng.normal(size=(n, 3)) * vel; s[:, 22:25] = rng.normal(size=(n, 3)) * 2 * vel
    s[:, 25:37] = rng.normal(size=(n, 12)) * 5 * vel
    return s


def stance_states(rng, n):
    """upright robots near the ground with feet at / slightly below the surface, moving: contact-rich"""
    from oracle_lib import STATE_DIM
    s = np.zeros((n, STATE_DIM))
    s[:, 0:2] = rng.uniform(-5, 5, size=(n, 2)); s[:, 2] = rng.uniform(0.24, 0.31, size=n)
    ang = rng.normal(size=(n, 3)) * 0.08
    s[:, 3] = 1.0; s[:, 4:7] = ang / 2
    s[:, 3:7] /= np.linalg.norm(s[:, 3:7], axis=1, keepdims=True)
    s[:, 7:19] = np.tile([0.0, -0.78, 1.57], 4) + rng.uniform(-0.25, 0.25, size=(n, 12))
    s[:, 19:22] = rng.normal(size=(n, 3)) * np.array([1.5, 0.3, 0.5]); s[:, 22:25] = rng.normal(size=(n, 3)) * 1.0
    s[:, 25:37] = rng.normal(size=(n, 12)) * 4
    return s
