"""`RaisimGymVecEnv`: the VecEnv adapter between the native vec-env and the RL code -- same public surface as the reference's
flex_gym/env/RaisimGymVecEnv.py (attributes `wrapper, num_obs, num_acts, _observation, _reward, _done, _extraInfo, rewards`;
`step / reset / reset_and_update_info`; the probe getters; window / video / curriculum pass-throughs; the `num_envs`,
`observation_space`, `action_space`, `extra_info_names` properties), without gym / stable_baselines and without O(N) Python per step:

  * `info` is a lazy sequence.  The reference materialises 6N one-key dicts per step (entry k holds extra-info j = k // N of env
    i = k % N, and entry i also gets env i's `episode`), of which only `info[i]['episode']` is ever read (ppo2.py:534-537).
    `LazyInfo` builds an entry only when it is indexed, with the same content.
  * episode return / length bookkeeping happens in the step kernel; `self.rewards` is the reference's list of per-env reward lists
    (RaisimGymVecEnv.py:21, 42-50) up to `TRACK_REWARDS_MAX_ENVS` environments (the reference's own scale: 200 envs) and an O(1)
    stand-in above it unless `track_rewards=True` is passed: N Python list appends per step are what dominates at N >= 16k.
  * numpy >= 1.24 spellings (`bool`, `np.inf`).
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import numpy as np


class Box:
    """Shape / bounds / dtype record standing in for gym.spaces.Box."""

    def __init__(self, low, high, dtype=np.float32):
        self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
        self.shape, self.dtype = self.low.shape, np.dtype(dtype)

    def __repr__(self):
        return f"Box{self.shape}"


class LazyInfo(Sequence):
    def __init__(self, names: List[str], extra: np.ndarray, episodes: Dict[int, Dict[str, Any]]):
        self._names, self._extra, self._episodes, self._n = names, extra, episodes, extra.shape[0]

    def __len__(self):
        return max(len(self._names), 1) * self._n

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        k = k + len(self) if k < 0 else k
        if not 0 <= k < len(self):
            raise IndexError(k)
        entry: Dict[str, Any] = {}
        if self._names:
            j, i = divmod(k, self._n)
            entry["extra_info"] = {self._names[j]: self._extra[i, j]}
        if k in self._episodes:
            entry["episode"] = self._episodes[k]
        return entry

    def copy(self):
        return LazyInfo(self._names, self._extra.copy(), dict(self._episodes))

    def episodes(self):
        """the `episode` records only, by env index"""
        return [self._episodes[i] for i in sorted(self._episodes)]


_UNSUPPORTED = ("render", "step_async", "step_wait", "get_attr", "set_attr", "env_method")
# probe getter -> (native method, row width as a function of the action count)
_PROBES = {
    "ReferenceState": ("ReferenceState", lambda a: 2 * a),
    "GetJointEffort": ("GetJointEffort", lambda a: a),
    "GetGeneralizedForce": ("GetGeneralizedForce", lambda a: a + 6),          # floating base: 6 + joints
    "GetInverseMassMatrix": ("GetInverseMassMatrix", lambda a: (a + 6) ** 2),
    "GetNonlinear": ("GetNonlinear", lambda a: a + 6),
    "GetSphereInfo": ("GetSphereInfo", lambda a: 4),
}
_PASSTHROUGH = {"close": "close", "start_recording_video": "startRecordingVideo", "stop_recording_video": "stopRecordingVideo",
                "curriculum_callback": "curriculumUpdate", "show_window": "showWindow", "hide_window": "hideWindow"}


TRACK_REWARDS_MAX_ENVS = 1024


class _UntrackedRewards(Sequence):
    """`rewards` above TRACK_REWARDS_MAX_ENVS: indexable like the reference's list of lists, every per-env history empty"""

    def __init__(self, n):
        self._n = n

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [[] for _ in range(*i.indices(self._n))]
        if not -self._n <= i < self._n:
            raise IndexError(i)
        return []


class RaisimGymVecEnv:
    def __init__(self, impl, track_rewards=None):
        self.wrapper = impl
        impl.init()
        n = impl.getNumOfEnvs()
        self.num_obs, self.num_acts = impl.getObDim(), impl.getActionDim()
        self._observation_space = Box(np.full(self.num_obs, -np.inf), np.full(self.num_obs, np.inf))
        self._action_space = Box(-np.ones(self.num_acts), np.ones(self.num_acts))
        self._extraInfoNames = impl.getExtraInfoNames()
        self._ep_ret, self._ep_len = np.zeros(n, np.float32), np.zeros(n, np.int32)
        try:   # the adapter owns these buffers: one page-locked block laid out like the native output block -> one DMA copy per step
            from . import _lib
            self._observation, self._reward, self._extraInfo, self._done = _lib.pinned_block(
                ((n, self.num_obs), np.float32), ((n,), np.float32), ((n, len(self._extraInfoNames)), np.float32), ((n,), np.bool_))
        except Exception:
            self._observation = np.zeros((n, self.num_obs), np.float32)
            self._reward = np.zeros(n, np.float32)
            self._done = np.zeros(n, bool)
            self._extraInfo = np.zeros((n, len(self._extraInfoNames)), np.float32)
        self._track = (n <= TRACK_REWARDS_MAX_ENVS) if track_rewards is None else bool(track_rewards)
        self.rewards = [[] for _ in range(n)] if self._track else _UntrackedRewards(n)

    # ------------------------------------------------------------------ hot path
    def step(self, action, visualize=False):
        native = self.wrapper.testStep if visualize else self.wrapper.step
        native(np.ascontiguousarray(action, np.float32), self._observation, self._reward, self._done, self._extraInfo)
        episodes: Dict[int, Dict[str, Any]] = {}
        if self._done.any():
            self.wrapper.lastEpisodeStats(self._ep_ret, self._ep_len)
            episodes = {int(i): {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i])} for i in np.flatnonzero(self._done)}
        if self._track:
            for i, r in enumerate(self._reward):
                self.rewards[i].append(r)
                if self._done[i]:
                    self.rewards[i].clear()
        return self._observation.copy(), self._reward.copy(), self._done.copy(), LazyInfo(self._extraInfoNames, self._extraInfo.copy(), episodes)

    def reset(self):
        self._reward[:] = 0
        self.wrapper.reset(self._observation)
        return self._observation.copy()

    def reset_and_update_info(self):
        # the device counters are cleared by the reset, so they are read first (the reference reads its Python lists afterwards)
        info = self._update_epi_info()
        return self.reset(), info

    def _update_epi_info(self):
        self.wrapper.runningEpisodeStats(self._ep_ret, self._ep_len, True)
        if self._track:
            for r in self.rewards:
                r.clear()
        return [{"episode": {"r": float(r), "l": int(l)}} for r, l in zip(self._ep_ret, self._ep_len)]

    def seed(self, seed=None):
        self.wrapper.setSeed(0 if seed is None else int(seed))      # the reference calls a non-existent wrapper.seed here

    # ------------------------------------------------------------------ probes and pass-throughs
    def _fetch(self, native_name: str, width: int):
        out = np.zeros((self.num_envs, width), np.float32)
        getattr(self.wrapper, native_name)(out)
        return out

    def OriginState(self):
        return self._fetch("OriginState", self.wrapper.GetOriginStateDim())

    def SetContactCoefficient(self, contact_coeff):
        self.wrapper.SetContactCoefficient(np.ascontiguousarray(contact_coeff, np.float32))

    def __getattr__(self, name):
        if name in _PROBES:
            native, width = _PROBES[name]
            return lambda: self._fetch(native, width(self.num_acts))
        if name in _PASSTHROUGH:
            return getattr(self.wrapper, _PASSTHROUGH[name])
        if name in _UNSUPPORTED:
            def _raise(*a, **k):
                raise RuntimeError('This method is not implemented')
            return _raise
        raise AttributeError(name)

    num_envs = property(lambda self: self.wrapper.getNumOfEnvs())
    observation_space = property(lambda self: self._observation_space)
    action_space = property(lambda self: self._action_space)
    extra_info_names = property(lambda self: self._extraInfoNames)


class fused_host_step:
    """One native call per control step of a host-side rollout loop (include/irrl_b200.h irrl_act_step): model.step -> clip ->
    env.step with page-locked host buffers owned here.  `obs` / `done` are inputs and outputs (Runner.run threads them through,
    ppo2.py:519-538); `action, clipped, value, neglogp, reward, extra` are the step's results.  `state` is a device tensor / pointer.

        fs = fused_host_step(env.wrapper, policy, state, stream_ptr)
        env.wrapper.reset(fs.obs)
        for t in range(T): fs(tick0 + t); mb_obs.append(fs.obs.copy()) ...
    """

    def __init__(self, wrapper, policy, state, stream_ptr=None, seed=None, env_offset=None, chunks=0, deterministic=False):
        import ctypes as C
        from . import _lib
        self._L, self._C = _lib.load(), C
        n = wrapper.getNumOfEnvs()
        self.action, self.clipped, self.value, self.neglogp, self.obs, self.reward, self.extra, self.done = _lib.pinned_block(
            ((n, 12), np.float32), ((n, 12), np.float32), ((n,), np.float32), ((n,), np.float32),
            ((n, 35), np.float32), ((n,), np.float32), ((n, 6), np.float32), ((n,), np.bool_))
        self._wrapper, self._policy, self._state = wrapper, policy, state          # keep the owners alive
        sp = state.data_ptr() if hasattr(state, "data_ptr") else int(state)
        p = lambda a: a.ctypes.data
        self._io = _lib.ActStepIO(obs=p(self.obs), done=p(self.done), state=sp, action=p(self.action), clipped=p(self.clipped), value=p(self.value),
                                  neglogp=p(self.neglogp), next_obs=p(self.obs), reward=p(self.reward), next_done=p(self.done), extra=p(self.extra))
        self._chunks, self._det = int(chunks), int(bool(deterministic))
        self.h2d_bytes = n * (35 * 4 + 1) if (chunks > 1 or (chunks == 0 and n >= 12288)) else n * (42 * 4 + 1)   # the un-chunked path moves the whole [obs | reward | extra | done] block
        self.d2h_bytes = n * (12 * 4 * 2 + 4 + 4 + 35 * 4 + 4 + 6 * 4 + 1)

    def __call__(self, tick):
        from . import _lib
        _lib.check(self._L.irrl_act_step(self._wrapper.handle, self._policy.handle, self._C.byref(self._io), self._det, int(tick) & 0xFFFFFFFF, self._chunks), "act_step")
