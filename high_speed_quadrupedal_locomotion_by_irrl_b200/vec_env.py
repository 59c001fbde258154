"""`RaisimGymVecEnv` -- the stable-baselines VecEnv adapter of the reference (flex_gym/env/RaisimGymVecEnv.py:6-189)
with the same attributes, methods, return values and in-place buffer semantics, minus the imports that do not exist
here (gym, stable_baselines) and minus the O(N) Python per step:

  * `info` is a lazy sequence: the reference builds 6N one-key dicts every step (RaisimGymVecEnv.py:35-40) of which only
    `info[i]['episode']` is ever consumed (ppo2.py:534-537).  Here entries materialise on access with the same content
    (entry k < 6N holds extra-info j = k // N of env i = k % N; entry i additionally carries env i's `episode`).
  * episode return / length bookkeeping (RaisimGymVecEnv.py:42-50) runs in the step kernel; `self.rewards` keeps the
    reference's list-of-lists view only when `track_rewards=True`.
  * numpy >= 1.24 names: `bool` / `np.inf` instead of the removed `np.bool` / `np.Inf`.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import numpy as np


class Box:
    """Minimal stand-in for gym.spaces.Box (shape, low, high, dtype)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return f"Box{self.shape}"


class LazyInfo(Sequence):
    """Materialises the reference's `info` list on demand."""

    def __init__(self, names: List[str], extra: np.ndarray, episodes: Dict[int, Dict[str, Any]]):
        self._names, self._extra, self._episodes = names, extra, episodes
        self._n = extra.shape[0]

    def __len__(self):
        return len(self._names) * self._n if self._names else self._n

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        if not 0 <= k < len(self):
            raise IndexError(k)
        d: Dict[str, Any] = {}
        if self._names:
            j, i = divmod(k, self._n)
            d["extra_info"] = {self._names[j]: self._extra[i, j]}
        if k in self._episodes:
            d["episode"] = self._episodes[k]
        return d

    def copy(self):
        return self

    def episodes(self):
        """only the entries that carry an `episode` key (what ppo2.py:534-537 scans for)"""
        return [self._episodes[i] for i in sorted(self._episodes)]


class RaisimGymVecEnv:
    def __init__(self, impl, track_rewards: bool = False):
        self.wrapper = impl
        self.wrapper.init()
        self.num_obs = self.wrapper.getObDim()
        self.num_acts = self.wrapper.getActionDim()
        self._observation_space = Box(np.ones(self.num_obs) * -np.inf, np.ones(self.num_obs) * np.inf, dtype=np.float32)
        self._action_space = Box(np.ones(self.num_acts) * -1., np.ones(self.num_acts) * 1., dtype=np.float32)
        self._observation = np.zeros([self.num_envs, self.num_obs], dtype=np.float32)
        self._reward = np.zeros(self.num_envs, dtype=np.float32)
        self._done = np.zeros((self.num_envs), dtype=bool)
        self._extraInfoNames = self.wrapper.getExtraInfoNames()
        self._extraInfo = np.zeros([self.num_envs, len(self._extraInfoNames)], dtype=np.float32)
        self._ep_ret = np.zeros(self.num_envs, dtype=np.float32)
        self._ep_len = np.zeros(self.num_envs, dtype=np.int32)
        # the wrapper owns these buffers (as in the reference): page-lock them once so step() DMAs straight into them
        try:
            from . import _lib
            for buf in (self._observation, self._reward, self._done, self._extraInfo):
                _lib.pin(buf)
        except Exception:
            pass
        self._track = track_rewards
        self.rewards = [[] for _ in range(self.num_envs)] if track_rewards else None

    def seed(self, seed=None):
        # the reference calls a non-existent wrapper.seed (RaisimGymVecEnv.py:24); setSeed is the bound name
        self.wrapper.setSeed(0 if seed is None else int(seed))

    def step(self, action, visualize=False):
        action = np.ascontiguousarray(action, dtype=np.float32)
        if not visualize:
            self.wrapper.step(action, self._observation, self._reward, self._done, self._extraInfo)
        else:
            self.wrapper.testStep(action, self._observation, self._reward, self._done, self._extraInfo)
        episodes: Dict[int, Dict[str, Any]] = {}
        if self._done.any():
            self.wrapper.lastEpisodeStats(self._ep_ret, self._ep_len)
            for i in np.flatnonzero(self._done):
                episodes[int(i)] = {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i])}
        if self._track:
            for i in range(self.num_envs):
                self.rewards[i].append(self._reward[i])
                if self._done[i]:
                    self.rewards[i].clear()
        info = LazyInfo(self._extraInfoNames, self._extraInfo.copy(), episodes)
        return self._observation.copy(), self._reward.copy(), self._done.copy(), info

    def _probe(self, fn, width):
        temp = np.zeros([self.num_envs, width], dtype=np.float32)
        fn(temp)
        return temp

    def OriginState(self):
        return self._probe(self.wrapper.OriginState, self.wrapper.GetOriginStateDim())

    def ReferenceState(self):
        return self._probe(self.wrapper.ReferenceState, self.num_acts * 2)

    def GetJointEffort(self):
        return self._probe(self.wrapper.GetJointEffort, self.num_acts)

    def GetGeneralizedForce(self):
        return self._probe(self.wrapper.GetGeneralizedForce, self.num_acts + 6)

    def GetInverseMassMatrix(self):
        return self._probe(self.wrapper.GetInverseMassMatrix, (self.num_acts + 6) * (self.num_acts + 6))

    def GetNonlinear(self):
        return self._probe(self.wrapper.GetNonlinear, self.num_acts + 6)

    def GetSphereInfo(self):
        return self._probe(self.wrapper.GetSphereInfo, 4)

    def SetContactCoefficient(self, contact_coeff):
        self.wrapper.SetContactCoefficient(np.ascontiguousarray(contact_coeff, dtype=np.float32))

    def reset(self):
        self._reward = np.zeros(self.num_envs, dtype=np.float32)
        self.wrapper.reset(self._observation)
        return self._observation.copy()

    def reset_and_update_info(self):
        # the reference resets first and reads the running episode statistics afterwards (RaisimGymVecEnv.py:100-101);
        # the device counters are cleared by reset, so they are read first here -- same values
        info = self._update_epi_info()
        return self.reset(), info

    def _update_epi_info(self):
        self.wrapper.runningEpisodeStats(self._ep_ret, self._ep_len, True)
        info = [{"episode": {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i])}} for i in range(self.num_envs)]
        if self._track:
            for r in self.rewards:
                r.clear()
        return info

    def render(self, mode='human'):
        raise RuntimeError('This method is not implemented')

    def close(self):
        self.wrapper.close()

    def start_recording_video(self, file_name):
        self.wrapper.startRecordingVideo(file_name)

    def stop_recording_video(self):
        self.wrapper.stopRecordingVideo()

    def curriculum_callback(self):
        self.wrapper.curriculumUpdate()

    def step_async(self):
        raise RuntimeError('This method is not implemented')

    def step_wait(self):
        raise RuntimeError('This method is not implemented')

    def get_attr(self, attr_name, indices=None):
        raise RuntimeError('This method is not implemented')

    def set_attr(self, attr_name, value, indices=None):
        raise RuntimeError('This method is not implemented')

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        raise RuntimeError('This method is not implemented')

    def show_window(self):
        self.wrapper.showWindow()

    def hide_window(self):
        self.wrapper.hideWindow()

    @property
    def num_envs(self):
        return self.wrapper.getNumOfEnvs()

    @property
    def observation_space(self):
        return self._observation_space

    @property
    def action_space(self):
        return self._action_space

    @property
    def extra_info_names(self):
        return self._extraInfoNames
