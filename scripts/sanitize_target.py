#!/usr/bin/env python
"""Target of scripts/r2_sanitize.sh (compute-sanitizer memcheck / racecheck / synccheck): every kernel of the hot path on small, awkward
shapes -- 300 environments (ragged last tile of the tcgen05 act kernel, partial last CTA of the step kernel), forced resets and trunk-box contacts (loop hand-over) inside the
step kernel, heightfield terrain, the meteor sphere, the 128-thread barrier variant of the step kernel, the fused host entry with chunks,
the GAE kernel and the learner kernels (tensor-core sequence-persistent BPTT, streaming projections / weight gradients) on ragged tiles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, dump_yaml
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy, PARAM_NAMES
from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import fused_host_step
from gpu_lib import Cuda
z = np.load(os.path.join(ROOT, "tests", "golden", "bp5_155_params.npz")); W = [z[k] for k in PARAM_NAMES]
n = 300
rng = np.random.default_rng(0)
for name, kw in (("flat + forced resets", dict()), ("stairs", dict(Terrain=True, terrain_kind="stairs")), ("meteor", dict(Crutial=True)),
                 ("obs filter", dict(ObsFilter=True))):
    c = Cuda(trot_cfg(num_envs=n, StochasticDynamics=True, ObsNoise=2.0, **kw))
    if "stairs" in name:
        pass
    pol = FusedLstmPolicy(W, n_env=n, seed=1)
    ob = c.reset(); st = np.zeros((n, 384), np.float32); done = np.zeros(n, bool); nd = 0
    for t in range(5):
        if t == 2:   # tip a third of the robots over: terminations + auto-resets inside the step kernel
            s = c.get_state(); s[::3, 3] = np.cos(0.6); s[::3, 4] = np.sin(0.6); s[::3, 5:7] = 0; c.set_state(s)
        if t == 3:   # lay another third on the belly: trunk-box contacts, i.e. the hot substep loop hands over to the complete one
            s = c.get_state(); s[1::3, 2] = 0.06; s[1::3, 7:19] = np.tile([0.0, -1.4, 2.6], 4).astype(np.float32); c.set_state(s)
        act, val, st, nlp = pol.step(ob, st, done)
        ob, rew, done, ex = c.step(np.clip(act, -1, 1)); nd += int(done.sum())
    print(f"{name}: 5 steps ok, {nd} auto-resets, reward mean {rew.mean():.3f}", flush=True)
os.environ["IRRL_STEP_BLK"] = "128"      # read once per process: the barrier variant is exercised through its own process run below
c = Cuda(trot_cfg(num_envs=n, StochasticDynamics=True, ObsNoise=2.0)); pol = FusedLstmPolicy(W, n_env=n, seed=1)
state = torch.zeros((n, 384), device="cuda:0")
for chunks in (1, 2):
    fs = fused_host_step(c.env, pol, state, chunks=chunks); c.env.reset(fs.obs)
    for t in range(3): fs(t)
print("fused host entry ok (1 and 2 chunks)", flush=True)
# device rollout + GAE + BPTT kernels through PPO2 on a tiny batch
from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import RaisimGymVecEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import PPO2
env = RaisimGymVecEnv(FlexibleGymEnv("", dump_yaml(trot_cfg(num_envs=76, StochasticDynamics=True, ObsNoise=2.0))))      # 76: ragged last tiles of every learner kernel (24- / 32- / 64- / 128-row tiles) and a multiple of 4, so that the bulk-copy (tcgen05) kernels are the ones that run
model = PPO2(env, policy_params=W, n_steps=12, noptepochs=1, learning_rate=1e-4, verbose=0)
h = model.learn(total_timesteps=76 * 12)
print("ppo iteration ok: loss", h[-1].get("policy_loss"), flush=True)
# the persistent tcgen05 learner kernels with several tiles per CTA (stage / slot / TMEM-buffer reuse), ragged last tiles
from high_speed_quadrupedal_locomotion_by_irrl_b200.lstm_seq import ProjRows, gram2_rows
g = torch.Generator(device="cuda:0"); g.manual_seed(5)
T, K, N = 48, 2, 520
obs_ = torch.randn(T, N, 35, device="cuda:0", generator=g); H0_ = torch.randn(T, K, N, 48, device="cuda:0", generator=g, requires_grad=True)
D_ = torch.randn(T, K, N, 192, device="cuda:0", generator=g); w0_ = torch.randn(K, 35, 192, device="cuda:0", generator=g); w1_ = torch.randn(K, 48, 192, device="cuda:0", generator=g, requires_grad=True)
ProjRows.apply(obs_, w0_); (ProjRows.apply(H0_, w1_) * D_).sum().backward(); h0_ = torch.randn(K, N, 48, device="cuda:0", generator=g); keep_ = (torch.rand(T, N, device="cuda:0", generator=g) > 0.1).float()
gram2_rows(obs_, H0_.detach(), h0_, keep_, D_); gram2_rows(H0_.detach(), H0_.detach(), h0_, keep_, D_)
torch.cuda.synchronize(); print("persistent tcgen05 learner kernels ok (multi-tile CTAs)", flush=True)
