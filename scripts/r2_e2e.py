#!/usr/bin/env python
"""e2e of irrl_act_step vs chunk count (GPU box): python scripts/r2_e2e.py 4096 16384"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import relaxation_cfg, dump_yaml
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy, PARAM_NAMES
from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import fused_host_step
z = np.load(os.path.join(ROOT, "tests", "golden", "bp5_155_params.npz")); W = [z[k] for k in PARAM_NAMES]
for n in [int(a) for a in sys.argv[1:]] or [4096, 16384]:
    env = FlexibleGymEnv("", dump_yaml(relaxation_cfg(num_envs=n, StochasticDynamics=True, ObsNoise=2.0))); env.init()
    pol = FusedLstmPolicy(W, n_env=n, seed=1); state = torch.zeros((n, 384), device="cuda:0")
    for chunks in (1, 2, 4, 8):
        fs = fused_host_step(env, pol, state, chunks=chunks); env.reset(fs.obs)
        for t in range(5): fs(t)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in range(100): fs(10 + t)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 100
        print(f"envs {n} chunks {chunks}: {dt * 1e6:.1f} us/step  {n / dt:.3e} env-steps/s", flush=True)
