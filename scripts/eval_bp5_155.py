#!/usr/bin/env python
"""Robot-level check against the reference's shipped results (SURVEY.md section 6 / BASELINE.md): the trained bp5_155 policy, driven
like run_bp_v5.py --test (Manual: True, bp5_test.yaml; the command is written into the network input, run_bp_v5.py:317-462) with a
forward-speed command ramped to 5 m/s, friction 0.8 (run_bp_v5.py:317-318).  Reference (RaiSim): mean vx 4.964 +- 0.071 m/s,
z 0.2732 +- 0.0016 m, stride 5.0 Hz (Exp_Raw_Data/body-center-2021-06-22-16-48-33.bin); eval.png: z 0.275 +- 0.004, vx error -0.066 +- 0.067.
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import test_cfg, dump_yaml
from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import RaisimGymVecEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy, PARAM_NAMES


def run(vx_cmd=5.0, mu=0.8, n=8, seconds=8.0, stand_height=0.28, verbose=False):
    cfg = test_cfg(num_envs=n, render=False, stand_height=stand_height)
    env = RaisimGymVecEnv(FlexibleGymEnv("", dump_yaml(cfg)))
    env.SetContactCoefficient(np.tile(np.array([[mu, 0.2, 0.01]], np.float32), (n, 1)))
    z = np.load(os.path.join(ROOT, "tests", "golden", "bp5_155_params.npz")); W = [z[k] for k in PARAM_NAMES]
    pol = FusedLstmPolicy(W, n_env=n)
    obs = env.reset(); state = np.zeros((n, 384), np.float32); done = np.zeros(n, bool)
    T = int(seconds / 0.002); v = 0.0
    rec = np.zeros((T, n, 41), np.float32); falls = 0
    for t in range(T):
        v = 0.999 * v + 0.001 * vx_cmd                       # the ramp of the reference's trot_ref_.csv dump
        obs[:, 0] = (v - 2.5) / 1.0; obs[:, 1] = 0.0; obs[:, 2] = 0.0      # bp5_config.py:19-55 scaling (obMean Vx/2, std 1)
        act, val, state, nlp = pol.step(obs, state, done, deterministic=True)
        obs, rew, done, info = env.step(np.clip(act, -1, 1))
        falls += int(done.sum())
        rec[t] = env.OriginState()
    half = rec[T // 2:]
    vx = half[:, :, 19]; zz = half[:, :, 2]
    # stride frequency from the pitch rate spectrum (body-frame y angular velocity ~ world for small attitude)
    def peak_hz(sig):
        sig = sig - sig.mean(); f = np.fft.rfftfreq(len(sig), 0.002); sp = np.abs(np.fft.rfft(sig)); return float(f[1:][np.argmax(sp[1:])])
    stride_hz = peak_hz(half[:, 0, 8])          # FR hip joint angle: one cycle per stride
    bounce_hz = peak_hz(half[:, 0, 21])         # vertical trunk velocity: two bounces per trot stride
    q = half[:, :, 3:7]; roll = np.arctan2(2 * (q[..., 0] * q[..., 1] + q[..., 2] * q[..., 3]), 1 - 2 * (q[..., 1] ** 2 + q[..., 2] ** 2))
    pitch = np.arcsin(np.clip(2 * (q[..., 0] * q[..., 2] - q[..., 3] * q[..., 1]), -1, 1))
    out = dict(cmd_vx=vx_cmd, mu=mu, falls=falls, vx_mean=float(vx.mean()), vx_std=float(vx.std()), z_mean=float(zz.mean()), z_std=float(zz.std()),
               roll_mean=float(roll.mean()), pitch_mean=float(pitch.mean()), stride_hz=stride_hz, bounce_hz=bounce_hz, final_cmd=float(v), cmd_mean_second_half=float(np.mean([vx_cmd * (1 - 0.999 ** (k + 1)) for k in range(T // 2, T)])))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--vx", type=float, default=5.0); ap.add_argument("--mu", type=float, default=0.8)
    ap.add_argument("--seconds", type=float, default=8.0); ap.add_argument("--stand", type=float, default=0.28)
    a = ap.parse_args()
    print(json.dumps(run(a.vx, a.mu, seconds=a.seconds, stand_height=a.stand)))
