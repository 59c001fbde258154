#!/usr/bin/env python
"""PPO iteration wall-time (BASELINE.json metric, second half): rollout of T steps on the device + GAE + noptepochs BPTT epochs.
   python scripts/ppo_timing.py --envs 4096 --steps 750 --epochs 10        (torchrun for N > 1 GPUs: envs are per GPU)"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, dump_yaml
from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import RaisimGymVecEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES
from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import PPO2
ap = argparse.ArgumentParser(); ap.add_argument("--envs", type=int, default=4096); ap.add_argument("--steps", type=int, default=750)
ap.add_argument("--epochs", type=int, default=10); ap.add_argument("--iters", type=int, default=2); ap.add_argument("--tf32", action="store_true"); a = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
z = np.load(os.path.join(ROOT, "tests", "golden", "bp5_155_params.npz")); W = [z[k] for k in PARAM_NAMES]
env = RaisimGymVecEnv(FlexibleGymEnv("", dump_yaml(trot_cfg(num_envs=a.envs, StochasticDynamics=True, ObsNoise=2.0)), device=local, env_offset=rank * a.envs))
model = PPO2(env, policy_params=W, n_steps=a.steps, noptepochs=a.epochs, learning_rate=1e-4, verbose=0, matmul_tf32=a.tf32)
hist = model.learn(total_timesteps=a.envs * a.steps * world * a.iters)
if rank == 0:
    h = hist[-1]
    print(json.dumps({"metric": "PPO iteration wall-time", "envs_per_gpu": a.envs, "n_gpus": world, "n_steps": a.steps, "noptepochs": a.epochs,
                      "iteration_s": h["iteration_s"], "rollout_s": h["rollout_s"], "update_s": h["update_s"], "fps": h["fps"],
                      "ep_reward_mean": h["ep_reward_mean"], "policy_loss": h["policy_loss"], "value_loss": h["value_loss"], "matmul_tf32": a.tf32}))
if world > 1:
    torch.distributed.destroy_process_group()
