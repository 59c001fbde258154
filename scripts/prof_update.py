import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, dump_yaml
from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import RaisimGymVecEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES
from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import PPO2
z = np.load(os.path.join(ROOT, "tests", "golden", "bp5_155_params.npz")); W = [z[k] for k in PARAM_NAMES]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
env = RaisimGymVecEnv(FlexibleGymEnv("", dump_yaml(trot_cfg(num_envs=N, StochasticDynamics=True, ObsNoise=2.0))))
model = PPO2(env, policy_params=W, n_steps=750, noptepochs=2, learning_rate=1e-4, verbose=0)
mb_states, _ = model._rollout()
model._update(mb_states, 1e-4, 0.2); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    model._update(mb_states, 1e-4, 0.2); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
