"""PPO2 with the two-tower LSTM policy on PyTorch, fed by the device-resident rollout (`irrl_rollout`).

Re-hosts flex_gym/algo/ppo2/ppo2.py (stable-baselines 2.8 / TF1) for the hot path's caller:
  * loss            ppo2.py:152-175  (clipped surrogate, clipped value loss, loss = pg - ent_coef*entropy + vf_coef*vf)
  * advantage norm  ppo2.py:262-263  over the WHOLE batch (all ranks: 3-scalar all-reduce)
  * optimiser       ppo2.py:191-197  clip_by_global_norm(0.5) then Adam(lr, eps=1e-5) in TF1's formulation
  * recurrent minibatching by env, `nminibatches` groups of envs, `noptepochs` passes  ppo2.py:381-404
  * Runner.run      ppo2.py:494-582  T steps of [model.step -> clip -> env.step], GAE(gamma, lam), forced reset; the LSTM
                    state and `dones` are NOT cleared by the forced reset
  * save / load     ppo2.py:452-476  parameter list in tf.trainable_variables() order (reference pkl files import)
Data parallel: one process per GPU, environments sharded, gradients all-reduced (sum / world) on a flat 70 741-float buffer.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .policy import PARAM_NAMES, PARAM_SHAPES, NUM_PARAMS, init_params, load_reference_params, save_params_npz

LOG2PI = math.log(2.0 * math.pi)


class LstmActorCritic(torch.nn.Module):
    """CustomLSTMPolicy (run_bp_v5.py:117-176) as a torch module; parameters named/ordered like the checkpoint."""

    def __init__(self, params: Optional[Sequence[np.ndarray]] = None):
        super().__init__()
        params = params if params is not None else init_params()
        for n, p in zip(PARAM_NAMES, params):
            self.register_parameter(n, torch.nn.Parameter(torch.as_tensor(np.asarray(p, np.float32)).clone()))

    def param_list(self) -> List[torch.nn.Parameter]:
        return [getattr(self, n) for n in PARAM_NAMES]

    def export_params(self) -> List[np.ndarray]:
        return [p.detach().cpu().numpy().copy() for p in self.param_list()]

    @staticmethod
    def _cell(x, c, h, wx, wh, b):
        z = x @ wx + h @ wh + b                       # SB lstm(): z = x.wx + h.wh + b ; i,f,o,g = split(z, 4)
        i, f, o, g = z.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        return c, h

    def forward_sequence(self, obs, masks, state):
        """obs [N,T,35], masks [N,T] (1.0 where the previous step was done), state [N,384] at t=0.
        Returns mean [N,T,12], value [N,T], final state [N,384]."""
        N, T, _ = obs.shape
        H = 48
        cs = [state[:, 0:H], state[:, 2 * H:3 * H], state[:, 4 * H:5 * H], state[:, 6 * H:7 * H]]      # [c0,h0,c1,h1]_pi | [..]_V
        hs = [state[:, H:2 * H], state[:, 3 * H:4 * H], state[:, 5 * H:6 * H], state[:, 7 * H:8 * H]]
        W = [(self.lstm_pi0_wx, self.lstm_pi0_wh, self.lstm_pi0_b), (self.lstm_pi1_wx, self.lstm_pi1_wh, self.lstm_pi1_b),
             (self.lstm_v0_wx, self.lstm_v0_wh, self.lstm_v0_b), (self.lstm_v1_wx, self.lstm_v1_wh, self.lstm_v1_b)]
        out_pi, out_v = [], []
        for t in range(T):
            keep = (1.0 - masks[:, t]).unsqueeze(1)
            x = obs[:, t]
            cs = [c * keep for c in cs]; hs = [h * keep for h in hs]
            cs[0], hs[0] = self._cell(x, cs[0], hs[0], *W[0])
            cs[1], hs[1] = self._cell(hs[0], cs[1], hs[1], *W[1])
            cs[2], hs[2] = self._cell(x, cs[2], hs[2], *W[2])
            cs[3], hs[3] = self._cell(hs[2], cs[3], hs[3], *W[3])
            out_pi.append(hs[1]); out_v.append(hs[3])
        lat_pi = torch.stack(out_pi, 1); lat_v = torch.stack(out_v, 1)
        mean = lat_pi @ self.pi_w + self.pi_b
        value = (lat_v @ self.vf_w + self.vf_b).squeeze(-1)
        new_state = torch.cat([cs[0], hs[0], cs[1], hs[1], cs[2], hs[2], cs[3], hs[3]], 1)
        return mean, value, new_state

    def features_time_major(self, obs, keep, state, fused: Optional[bool] = None):
        """obs [T,N,35] (time-major, as the device rollout stores it), keep [T,N] = 1 - mask, state [N,384] at t = 0.
        Returns the top-layer activations of both towers H1 [T,2,N,48] and whether the tensor-core learner kernels served them.
        On CUDA the recurrence runs through the fused BPTT path (lstm_seq.LstmLayerFused / LstmLayerSeqPersistent)."""
        from .lstm_seq import LstmLayerSeq, LstmLayerSeqPersistent, LstmLayerFused, fused_layer_ok, lstm_layer_reference, _with_bias, proj_rows
        T, N, _ = obs.shape
        H = 48
        fused = obs.is_cuda if fused is None else fused
        layer = lstm_layer_reference if not fused else (_with_bias(LstmLayerSeq.apply) if fused == "stepwise" else LstmLayerSeqPersistent.apply)
        st = state.view(N, 2, 2, 2, H)                                  # [env, tower, layer, (c,h), unit]
        c0 = st[:, :, 0, 0].transpose(0, 1).contiguous(); h0 = st[:, :, 0, 1].transpose(0, 1).contiguous()
        c1 = st[:, :, 1, 0].transpose(0, 1).contiguous(); h1 = st[:, :, 1, 1].transpose(0, 1).contiguous()
        wx0 = torch.stack([self.lstm_pi0_wx, self.lstm_v0_wx]); wh0 = torch.stack([self.lstm_pi0_wh, self.lstm_v0_wh]); b0 = torch.stack([self.lstm_pi0_b, self.lstm_v0_b])
        wx1 = torch.stack([self.lstm_pi1_wx, self.lstm_v1_wx]); wh1 = torch.stack([self.lstm_pi1_wh, self.lstm_v1_wh]); b1 = torch.stack([self.lstm_pi1_b, self.lstm_v1_b])
        own = fused is True                                                          # the persistent path: projections on the streaming tensor-core kernels too
        if own and fused_layer_ok(obs, wx0) and wx1.shape[1] == 48:                  # whole layers as one autograd node each (lstm_seq.LstmLayerFused)
            H0 = LstmLayerFused.apply(obs, wx0, wh0, b0, c0, h0, keep)
            H1 = LstmLayerFused.apply(H0, wx1, wh1, b1, c1, h1, keep)
        else:
            xw0 = proj_rows(obs, wx0) if own else torch.matmul(obs.unsqueeze(1), wx0)    # [T,2,N,192]: all input projections in one pass (bias added by the layer)
            H0, _, _ = layer(xw0, wh0, b0, c0, h0, keep)
            xw1 = proj_rows(H0, wx1) if own else torch.matmul(H0, wx1)
            H1, _, _ = layer(xw1, wh1, b1, c1, h1, keep)
        return H1, own

    def forward_time_major(self, obs, keep, state, fused: Optional[bool] = None):
        """Returns mean [T,N,12], value [T,N] (features_time_major + the two heads)."""
        H1, own = self.features_time_major(obs, keep, state, fused)
        if own:   # both heads as ONE batched product on the [T,2,N,48] tensor (vf_w zero-padded to 12 columns): no strided tower slices, no
            #       zero-filled select_backward copies of 2.4 GB each in the backward pass
            head_w = torch.stack([self.pi_w, F.pad(self.vf_w, (0, self.pi_w.shape[1] - self.vf_w.shape[1]))])
            out = torch.matmul(H1, head_w)                                           # [T,2,N,12]
            return out[:, 0] + self.pi_b, out[:, 1, :, 0] + self.vf_b
        mean = H1[:, 0] @ self.pi_w + self.pi_b
        value = (H1[:, 1] @ self.vf_w + self.vf_b).squeeze(-1)
        return mean, value

    def neglogp(self, mean, actions):
        logstd = self.pi_logstd.reshape(1, 1, -1)
        return 0.5 * (((actions - mean) / logstd.exp()) ** 2).sum(-1) + 0.5 * LOG2PI * actions.shape[-1] + logstd.sum()

    def entropy(self):
        return (self.pi_logstd + 0.5 * (LOG2PI + 1.0)).sum()


def ppo_loss(model: LstmActorCritic, obs, masks, state, actions, advs, returns, old_values, old_neglogp, cliprange, ent_coef, vf_coef, time_major=False):
    """ppo2.py:152-175.  Env-major [N,T,...] tensors (swap_and_flatten order) or, with time_major=True, [T,N,...] ones
    (the device rollout's native layout, served by the fused BPTT path); advs already normalised."""
    if time_major and obs.is_cuda and os.environ.get("IRRL_LEARNER_HEADS", "") != "torch" and os.environ.get("IRRL_LEARNER_GEMM", "") != "cublas":
        # heads + loss + their gradients as ONE kernel over the top-layer activations (lstm_seq.HeadLossFused)
        from .lstm_seq import HeadLossFused
        H1, own = model.features_time_major(obs, 1.0 - masks, state)
        if own:
            main, st = HeadLossFused.apply(H1, model.pi_w, model.pi_b, model.vf_w, model.vf_b, model.pi_logstd, actions, advs, returns, old_values, old_neglogp,
                                           float(cliprange), float(vf_coef))
            entropy = model.entropy()
            return main - entropy * ent_coef, dict(policy_loss=st[0], value_loss=st[1], policy_entropy=entropy.detach(), approxkl=st[2], clipfrac=st[3])
    if time_major:
        mean, vpred = model.forward_time_major(obs, 1.0 - masks, state)
    else:
        mean, vpred, _ = model.forward_sequence(obs, masks, state)
    neglogpac = model.neglogp(mean, actions)
    entropy = model.entropy()
    vpredclipped = old_values + torch.clamp(vpred - old_values, -cliprange, cliprange)
    vf_loss = 0.5 * torch.maximum((vpred - returns) ** 2, (vpredclipped - returns) ** 2).mean()
    ratio = torch.exp(old_neglogp - neglogpac)
    pg_loss = torch.maximum(-advs * ratio, -advs * torch.clamp(ratio, 1.0 - cliprange, 1.0 + cliprange)).mean()
    approxkl = 0.5 * ((neglogpac - old_neglogp) ** 2).mean()
    clipfrac = ((ratio - 1.0).abs() > cliprange).float().mean()
    loss = pg_loss - entropy * ent_coef + vf_loss * vf_coef
    return loss, dict(policy_loss=pg_loss.detach(), value_loss=vf_loss.detach(), policy_entropy=entropy.detach(), approxkl=approxkl.detach(), clipfrac=clipfrac.detach())


class TF1Adam:
    """tf.train.AdamOptimizer(lr, epsilon=1e-5) semantics (ppo2.py:195-196): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t * m / (sqrt(v) + eps) -- epsilon is NOT bias-corrected, unlike torch.optim.Adam."""

    def __init__(self, params: Sequence[torch.Tensor], eps=1e-5, b1=0.9, b2=0.999):
        self.params = list(params); self.eps, self.b1, self.b2, self.t = eps, b1, b2, 0
        self.m = [torch.zeros_like(p) for p in self.params]; self.v = [torch.zeros_like(p) for p in self.params]

    @torch.no_grad()
    def step(self, grads: Sequence[torch.Tensor], lr: float):
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        torch._foreach_mul_(self.m, self.b1); torch._foreach_add_(self.m, grads, alpha=1.0 - self.b1)
        torch._foreach_mul_(self.v, self.b2); torch._foreach_addcmul_(self.v, grads, grads, value=1.0 - self.b2)
        denom = torch._foreach_sqrt(self.v); torch._foreach_add_(denom, self.eps)
        torch._foreach_addcdiv_(self.params, self.m, denom, value=-lr_t)


def clip_by_global_norm(grads: Sequence[torch.Tensor], max_norm: float):
    """tf.clip_by_global_norm (ppo2.py:191-193)."""
    gn = torch.sqrt(sum((g.float() ** 2).sum() for g in grads))
    scale = max_norm / torch.maximum(gn, torch.as_tensor(max_norm, device=gn.device, dtype=gn.dtype))
    return [g * scale for g in grads], gn


def dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def global_mean_std(x: torch.Tensor):
    """mean / std (population, like numpy's advs.std()) over all ranks: all-reduce of (sum, sumsq, count)  ppo2.py:263."""
    dist, _, world = dist_info()
    s = torch.stack([x.sum().double(), (x.double() ** 2).sum(), torch.as_tensor(float(x.numel()), device=x.device, dtype=torch.float64)])
    if world > 1:
        dist.all_reduce(s)
    mean = s[0] / s[2]
    var = torch.clamp(s[1] / s[2] - mean * mean, min=0.0)
    return mean.to(x.dtype), torch.sqrt(var).to(x.dtype)


ALLREDUCE_EVENTS = None   # set to a list to collect (start, end) CUDA events of every gradient all-reduce


def allreduce_grads(grads: Sequence[torch.Tensor]):
    """one flat buffer (70 741 floats = 283 KB), sum over ranks / world."""
    dist, _, world = dist_info()
    if world == 1:
        return list(grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    if ALLREDUCE_EVENTS is not None:      # bench.py: device time of the one collective of the workload (CUDA events around the NCCL call)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_reduce(flat); e1.record(); ALLREDUCE_EVENTS.append((e0, e1))
    else:
        dist.all_reduce(flat)
    flat /= world
    out, o = [], 0
    for g in grads:
        out.append(flat[o:o + g.numel()].view_as(g)); o += g.numel()
    return out


class PPO2:
    """Constructor arguments as used by run_bp_v5.py:227-242."""

    def __init__(self, env, policy_params: Optional[Sequence[np.ndarray]] = None, gamma=0.99, n_steps=750, ent_coef=0.0, learning_rate=1e-3,
                 vf_coef=0.5, max_grad_norm=0.5, lam=0.998, nminibatches=1, noptepochs=10, cliprange=0.2, verbose=1, device: Optional[int] = None,
                 seed: int = 0, matmul_tf32: bool = False):
        from .policy import FusedLstmPolicy
        # opt-in: run the learner's cuBLAS GEMMs (input projections, weight gradients) as single-pass TF32 tensor-core products.
        # Default False = fp32 like the reference's TF1 graph; rollout kernels are unaffected (they are fp32-accurate either way).
        self.matmul_tf32 = bool(matmul_tf32)
        self.env = env                                     # RaisimGymVecEnv over FlexibleGymEnv
        self.n_envs = env.num_envs
        self.gamma, self.lam, self.n_steps, self.ent_coef, self.vf_coef = gamma, lam, int(n_steps), ent_coef, vf_coef
        self.learning_rate, self.max_grad_norm, self.nminibatches, self.noptepochs, self.cliprange, self.verbose = learning_rate, max_grad_norm, nminibatches, noptepochs, cliprange, verbose
        assert self.n_envs % nminibatches == 0, "For recurrent policies, the number of environments run in parallel should be a multiple of nminibatches."
        self.device_index = env.wrapper.device if device is None else device
        self.dev = torch.device("cuda", self.device_index)
        self.model = LstmActorCritic(policy_params).to(self.dev)
        self.optim = TF1Adam(self.model.param_list())
        _, rank, world = dist_info()
        self.rank, self.world = rank, world
        if world > 1:   # equal shards only: gradients are averaged as sum / world of per-rank means and n_batch = n_envs * n_steps * world
            import torch.distributed as _d
            sizes = [None] * world; _d.all_gather_object(sizes, int(self.n_envs))
            if len(set(sizes)) != 1:
                raise ValueError(f"PPO2 needs the same number of environments on every rank, got {sizes}")
        self.act_model = FusedLstmPolicy(self.model.export_params(), n_env=self.n_envs, device=self.device_index, seed=seed, env_offset=rank * self.n_envs)
        self.num_timesteps = 0
        # Runner state (AbstractEnvRunner.__init__): obs = env.reset(), states = zeros, dones = False
        self._stream = torch.cuda.Stream(self.dev)
        with torch.cuda.device(self.dev):
            self.env.wrapper.setStream(self._stream.cuda_stream)
            self.cur_obs = torch.zeros((self.n_envs, 35), device=self.dev)
            self.env.wrapper.reset(self.cur_obs)
            self.cur_done = torch.zeros((self.n_envs,), device=self.dev, dtype=torch.uint8)
            self.state = torch.zeros((self.n_envs, 384), device=self.dev)
        self._buf = None

    # ---- Runner.run (ppo2.py:494-582) on the device
    def _rollout(self):
        from . import _lib
        L = _lib.load()
        T, N, dev = self.n_steps, self.n_envs, self.dev
        if self._buf is None:
            f = dict(device=dev, dtype=torch.float32)
            self._buf = dict(obs=torch.empty((T, N, 35), **f), actions=torch.empty((T, N, 12), **f), values=torch.empty((T, N), **f), neglogps=torch.empty((T, N), **f),
                             rewards=torch.empty((T, N), **f), dones=torch.empty((T, N), device=dev, dtype=torch.uint8), ep_return=torch.zeros((T, N), **f),
                             ep_length=torch.zeros((T, N), device=dev, dtype=torch.int32), adv=torch.empty((T, N), **f), ret=torch.empty((T, N), **f))
        b = self._buf
        mb_states = self.state.clone()
        rb = _lib.RolloutBuffers(obs=b["obs"].data_ptr(), actions=b["actions"].data_ptr(), values=b["values"].data_ptr(), neglogps=b["neglogps"].data_ptr(),
                                 rewards=b["rewards"].data_ptr(), dones=b["dones"].data_ptr(), cur_obs=self.cur_obs.data_ptr(), cur_done=self.cur_done.data_ptr(),
                                 state=self.state.data_ptr(), ep_return=b["ep_return"].data_ptr(), ep_length=b["ep_length"].data_ptr())
        with torch.cuda.stream(self._stream):
            self.act_model.set_params(self.model.export_params())
            _lib.check(L.irrl_rollout(self.env.wrapper.handle, self.act_model.handle, T, C.byref(rb), 0), "rollout")
            # last_values = model.value(obs, states, dones)  (ppo2.py:552): a deterministic act on a scratch copy of the state
            scratch = self.state.clone()
            lv = torch.empty((N,), device=dev); dummy_a = torch.empty((N, 12), device=dev); dummy_n = torch.empty((N,), device=dev)
            _lib.check(L.irrl_policy_act(self.act_model.handle, C.c_void_p(self._stream.cuda_stream), N, C.c_void_p(self.cur_obs.data_ptr()), C.c_void_p(self.cur_done.data_ptr()),
                                         C.c_void_p(scratch.data_ptr()), C.c_void_p(dummy_a.data_ptr()), None, C.c_void_p(lv.data_ptr()), C.c_void_p(dummy_n.data_ptr()), 1, 0, 0, 0), "value")
            _lib.check(L.irrl_gae(C.c_void_p(self._stream.cuda_stream), T, N, C.c_void_p(b["rewards"].data_ptr()), C.c_void_p(b["values"].data_ptr()), C.c_void_p(b["dones"].data_ptr()),
                                  C.c_void_p(lv.data_ptr()), C.c_void_p(self.cur_done.data_ptr()), C.c_float(self.gamma), C.c_float(self.lam), C.c_void_p(b["adv"].data_ptr()), C.c_void_p(b["ret"].data_ptr())), "gae")
            # episode infos: finished episodes inside the rollout + the running ones closed by the forced reset (ppo2.py:534-537, 577-580)
            fin = b["ep_length"] > 0
            ep_r = b["ep_return"][fin]; ep_l = b["ep_length"][fin]
            run_r = torch.empty((N,), device=dev); run_l = torch.empty((N,), device=dev, dtype=torch.int32)
            self.env.wrapper.runningEpisodeStats(run_r, run_l, True)
            self.env.wrapper.reset(self.cur_obs)          # "resetting environments, added by Jemin" ppo2.py:577; states/dones kept (quirk 10)
        self._stream.synchronize()
        ep_infos = dict(r=torch.cat([ep_r, run_r]).cpu().numpy(), l=torch.cat([ep_l, run_l]).cpu().numpy())
        return mb_states, ep_infos

    def _update(self, mb_states, lr, cliprange):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.matmul_tf32
        try:
            return self._update_impl(mb_states, lr, cliprange)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    def _update_impl(self, mb_states, lr, cliprange):
        b = self._buf
        N, T = self.n_envs, self.n_steps
        # the rollout buffers are time-major [T,N,...]; an env minibatch is a gather along dim 1 (the reference flattens env-major,
        # ppo2.py:572-574 / 385-386 -- same samples, the loss is a mean over them)
        obs, actions, returns, values, neglogps = b["obs"], b["actions"], b["ret"], b["values"], b["neglogps"]
        masks = b["dones"].float()
        envs_per_batch = N // self.nminibatches
        gen = torch.Generator(device="cpu"); gen.manual_seed(self.num_timesteps + 17 * self.rank)
        stats: List[Dict[str, torch.Tensor]] = []
        full = self.nminibatches == 1
        for epoch in range(self.noptepochs):
            perm = None if full else torch.randperm(N, generator=gen).to(self.dev)        # ppo2.py:388 shuffle env indices (irrelevant with one minibatch)
            for start in range(0, N, envs_per_batch):
                idx = None if full else perm[start:start + envs_per_batch]
                sel = (lambda x: x) if full else (lambda x: x.index_select(1, idx))        # with one minibatch the order is irrelevant
                advs = sel(returns) - sel(values)                                         # ppo2.py:262
                mean, std = global_mean_std(advs)
                advs = (advs - mean) / (std + 1e-8)                                       # ppo2.py:263
                loss, st = ppo_loss(self.model, sel(obs), sel(masks), mb_states if full else mb_states[idx], sel(actions), advs, sel(returns), sel(values),
                                    sel(neglogps), cliprange, self.ent_coef, self.vf_coef, time_major=True)
                grads = torch.autograd.grad(loss, self.model.param_list(), allow_unused=True)
                grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, self.model.param_list())]   # the q head is unused by the loss
                grads = allreduce_grads(grads)
                if self.max_grad_norm is not None:
                    grads, _ = clip_by_global_norm(grads, self.max_grad_norm)
                self.optim.step(grads, lr)
                stats.append(st)
        return {k: float(torch.stack([s[k] for s in stats]).mean().item()) for k in stats[0]}

    def learn(self, total_timesteps, log_interval=1, callback=None, save_every=100, log_dir: Optional[str] = None):
        n_batch = self.n_envs * self.n_steps * self.world
        nupdates = max(total_timesteps // n_batch, 1)
        t_first = time.time()
        history = []
        for update in range(1, nupdates + 1):
            t0 = time.time()
            frac = 1.0 - (update - 1.0) / nupdates
            lr = self.learning_rate(frac) if callable(self.learning_rate) else self.learning_rate
            clip = self.cliprange(frac) if callable(self.cliprange) else self.cliprange
            torch.cuda.nvtx.range_push("ppo2.rollout")
            mb_states, ep_infos = self._rollout()
            torch.cuda.nvtx.range_pop()
            t1 = time.time()
            torch.cuda.nvtx.range_push("ppo2.update")
            loss_vals = self._update(mb_states, lr, clip)
            torch.cuda.nvtx.range_pop()
            torch.cuda.synchronize(self.dev)
            t2 = time.time()
            self.num_timesteps += n_batch
            rec = dict(nupdates=update, total_timesteps=self.num_timesteps, fps=int(n_batch / (t2 - t0)), rollout_s=t1 - t0, update_s=t2 - t1, iteration_s=t2 - t0,
                       ep_reward_mean=float(np.mean(ep_infos["r"])) if len(ep_infos["r"]) else float("nan"),
                       ep_len_mean=float(np.mean(ep_infos["l"])) if len(ep_infos["l"]) else float("nan"), time_elapsed=t0 - t_first, **loss_vals)
            history.append(rec)
            if self.verbose >= 1 and self.rank == 0 and (update % log_interval == 0 or update == 1):
                print(" | ".join(f"{k} {v:.4g}" if isinstance(v, float) else f"{k} {v}" for k, v in rec.items()), flush=True)
            if log_dir and self.rank == 0 and (update % save_every == 1):
                self.save(f"{log_dir}_Iteration_{update - 1}")
            if callback is not None and callback(locals(), globals()) is False:
                break
        return history

    # ---- checkpoints (ppo2.py:452-476)
    def save(self, save_path: str):
        """PPO2.save (ppo2.py:452-476): `<save_path>.pkl` in the reference's `(data, params)` format (policy.save_reference_pkl), readable
        by the reference's PPO2.load / CustomerLstmNN; a path ending in .npz keeps the plain numpy archive."""
        if not save_path.endswith(".npz"):
            from .policy import save_reference_pkl
            return save_reference_pkl(save_path, self.model.export_params(), gamma=self.gamma, n_steps=self.n_steps, vf_coef=self.vf_coef, ent_coef=self.ent_coef,
                                      max_grad_norm=self.max_grad_norm, learning_rate=self.learning_rate if not callable(self.learning_rate) else self.learning_rate(1.0),
                                      lam=self.lam, nminibatches=self.nminibatches, noptepochs=self.noptepochs,
                                      cliprange=self.cliprange if not callable(self.cliprange) else self.cliprange(1.0), verbose=self.verbose, n_envs=self.n_envs * self.world)
        save_params_npz(save_path if save_path.endswith(".npz") else save_path + ".npz", self.model.export_params(), gamma=self.gamma, n_steps=self.n_steps,
                        vf_coef=self.vf_coef, ent_coef=self.ent_coef, max_grad_norm=self.max_grad_norm, lam=self.lam, nminibatches=self.nminibatches,
                        noptepochs=self.noptepochs, cliprange=self.cliprange if not callable(self.cliprange) else -1.0, n_envs=self.n_envs)

    @classmethod
    def load(cls, load_path: str, env, **kwargs):
        """`PPO2.load(path)` then `model.env = env` in the reference (run_bp_v5.py:245-248); optimiser state is not part of
        the reference checkpoint either."""
        return cls(env, policy_params=load_reference_params(load_path), **kwargs)

    def export_pi_csv(self, directory: str):
        """CustomerLstmNN.save_model (CustomerLstmNN.py:203-224): the pi tower as 8 CSV files with 6 decimals."""
        import os
        os.makedirs(directory, exist_ok=True)
        P = dict(zip(PARAM_NAMES, self.model.export_params()))
        for fn, key in (("lstm_wh0", "lstm_pi0_wh"), ("lstm_wh1", "lstm_pi1_wh"), ("lstm_wx0", "lstm_pi0_wx"), ("lstm_wx1", "lstm_pi1_wx"),
                        ("lstm_b0", "lstm_pi0_b"), ("lstm_b1", "lstm_pi1_b"), ("pi_w", "pi_w"), ("pi_b", "pi_b")):
            np.savetxt(os.path.join(directory, fn + ".csv"), P[key], delimiter=",", fmt="%.6f")
