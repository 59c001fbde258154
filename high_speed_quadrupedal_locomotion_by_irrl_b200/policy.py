"""Two-tower 2x48 LSTM actor-critic of the reference (`CustomLSTMPolicy`, run_bp_v5.py:117-193) on PyTorch, with the
rollout-time `step()` served by the fused CUDA act kernel through the C ABI.

Parameter order/shape = the reference checkpoint (ppo2.py:452-476): lstm_pi0{wx,wh,b},
lstm_pi1, lstm_v0, lstm_v1, vf{w,b}, pi{w,b}, pi/logstd, q{w,b} -- 19 arrays, 70 741 floats, so `bp5_155.pkl`
parameters load verbatim (`load_reference_params`).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib

PARAM_NAMES = ["lstm_pi0_wx", "lstm_pi0_wh", "lstm_pi0_b", "lstm_pi1_wx", "lstm_pi1_wh", "lstm_pi1_b",
               "lstm_v0_wx", "lstm_v0_wh", "lstm_v0_b", "lstm_v1_wx", "lstm_v1_wh", "lstm_v1_b",
               "vf_w", "vf_b", "pi_w", "pi_b", "pi_logstd", "q_w", "q_b"]
PARAM_SHAPES = [(35, 192), (48, 192), (192,), (48, 192), (48, 192), (192,), (35, 192), (48, 192), (192,), (48, 192),
                (48, 192), (192,), (48, 1), (1,), (48, 12), (12,), (1, 12), (48, 12), (12,)]
NUM_PARAMS = 70741
N_LSTM = [48, 48]   # run_bp_v5.py:111
STATE_DIM = 384


def flatten_params(params: Sequence[np.ndarray]) -> np.ndarray:
    assert len(params) == 19, "expected the 19 arrays of the reference checkpoint"
    for p, s in zip(params, PARAM_SHAPES):
        assert tuple(np.shape(p)) == s, (np.shape(p), s)
    flat = np.concatenate([np.asarray(p, np.float32).ravel() for p in params])
    assert flat.size == NUM_PARAMS
    return flat


def unflatten_params(flat: np.ndarray) -> List[np.ndarray]:
    out, o = [], 0
    for s in PARAM_SHAPES:
        n = int(np.prod(s))
        out.append(np.asarray(flat[o:o + n], np.float32).reshape(s).copy())
        o += n
    return out


def init_params(rng: Optional[np.random.Generator] = None) -> List[np.ndarray]:
    """Random initialisation in the style of stable-baselines' a2c.utils (orthogonal weights, zero biases; pi head
    scale 0.01, logstd 0) -- used when no checkpoint is given ("random-init weights of that architecture")."""
    rng = rng or np.random.default_rng(0)

    def ortho(shape, scale=1.0):
        a = rng.normal(size=shape)
        u, _, vt = np.linalg.svd(a, full_matrices=False)
        q = u if u.shape == shape else vt
        return (scale * q).astype(np.float32)

    out = []
    for s, name in zip(PARAM_SHAPES, PARAM_NAMES):
        if name.endswith("_b") or name == "pi_logstd":
            out.append(np.zeros(s, np.float32))
        elif name == "pi_w" or name == "q_w":
            out.append(ortho(s, 0.01))
        else:
            out.append(ortho(s, 1.0))
    return out


def load_reference_params(path: str) -> List[np.ndarray]:
    """Reads the parameter list of a reference checkpoint (cloudpickle `(data, params)`, ppo2.py:452-476) or of a
    `.npz` holding PARAM_NAMES.  The pickle is read with a stub unpickler: TF / stable-baselines classes referenced
    by the stream are replaced by inert placeholders, only the numpy arrays are kept."""
    if path.endswith(".npz"):
        z = np.load(path)
        return [np.asarray(z[k], np.float32) for k in PARAM_NAMES]
    import io
    import pickle

    class _Dummy:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Dummy()
        def __setstate__(self, s): self.__dict__["_state"] = s
        def __getattr__(self, n):
            if n.startswith("__"):
                raise AttributeError(n)
            return _Dummy()
        def __setitem__(self, k, v): pass
        def append(self, *a): pass
        def extend(self, *a): pass
        def update(self, *a, **k): pass

    class _U(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] in ("numpy", "builtins", "collections", "_codecs", "copyreg") and name not in ("eval", "exec", "__import__"):
                try:
                    return super().find_class(module, name)
                except Exception:
                    pass
            return _Dummy

    with open(path, "rb") as f:
        data, params = _U(io.BytesIO(f.read())).load()
    return [np.asarray(p, np.float32) for p in params]


def save_params_npz(path: str, params: Sequence[np.ndarray], **meta) -> None:
    np.savez(path, **{k: np.asarray(v, np.float32) for k, v in zip(PARAM_NAMES, params)}, **{f"meta_{k}": v for k, v in meta.items()})


#: tf.trainable_variables() names of the reference graph in checkpoint order (run_bp_v5.py:143-176; CustomerLstmNN.py:40-70 indexes by them)
TF_VARIABLE_NAMES = [f"model/lstm_{t}{i}/{w}:0" for t in ("pi", "v") for i in (0, 1) for w in ("wx", "wh", "b")] + \
                    ["model/vf/w:0", "model/vf/b:0", "model/pi/w:0", "model/pi/b:0", "model/pi/logstd:0", "model/q/w:0", "model/q/b:0"]


def save_reference_pkl(path: str, params: Sequence[np.ndarray], **hyper) -> str:
    """Writes a checkpoint in the reference's format (PPO2.save, ppo2.py:452-476 -> stable-baselines `_save_to_file`): one pickle
    stream holding the tuple `(data, params)` -- `data` the 17 constructor fields of ppo2.py:453-471, `params` the 19 arrays in
    tf.trainable_variables() order -- so that `PPO2.load(path)` / `run_bp_v5.py --load` and `CustomerLstmNN(path)`
    (CustomerLstmNN.py:27-60) read a model trained here.

    Three entries of `data` are objects of packages that do not exist in this image (stable-baselines policy class, two
    gym.spaces.Box).  They are written as pickle GLOBAL references by name, resolved by the *reading* process:
    `__main__.CustomLSTMPolicy` (the class run_bp_v5.py defines at line 117 -- the reference's own file stores that class by value
    with cloudpickle, which needs TensorFlow in the writer) and `gym.spaces.box.Box` with (low, high, shape, dtype) state."""
    import pickle
    import sys
    import types
    assert len(params) == 19
    params = [np.ascontiguousarray(p, np.float32).reshape(s) for p, s in zip(params, PARAM_SHAPES)]
    fake = {}

    def fake_class(module, name):
        if module not in sys.modules:
            parts = module.split(".")
            for i in range(1, len(parts) + 1):
                m = ".".join(parts[:i])
                if m not in sys.modules:
                    fake[m] = sys.modules[m] = types.ModuleType(m)
        cls = type(name, (), {"__module__": module})
        cls.__qualname__ = name
        had = getattr(sys.modules[module], name, None)
        setattr(sys.modules[module], name, cls)
        return cls, had

    policy_cls, had_policy = fake_class("__main__", "CustomLSTMPolicy")
    box_cls, had_box = fake_class("gym.spaces.box", "Box")

    def box(low, high, n):
        b = box_cls.__new__(box_cls)
        b.__dict__.update(low=np.full(n, low, np.float32), high=np.full(n, high, np.float32), shape=(n,), dtype=np.dtype(np.float32), np_random=None)
        return b

    data = dict(gamma=0.99, n_steps=750, vf_coef=0.5, ent_coef=0.0, max_grad_norm=0.5, learning_rate=1e-4, lam=0.998, nminibatches=1, noptepochs=10,
                cliprange=0.2, verbose=1, policy=policy_cls, observation_space=box(-np.inf, np.inf, 35), action_space=box(-1.0, 1.0, 12), n_envs=200,
                _vectorize_action=False, policy_kwargs=dict(n_lstm=list(N_LSTM)))
    for k, v in hyper.items():
        if k not in data:
            raise KeyError(f"{k} is not a field of the reference checkpoint (ppo2.py:453-471)")
        data[k] = v
    try:
        blob = pickle.dumps((data, params), protocol=2)
    finally:
        for name, cls, had, module in (("CustomLSTMPolicy", policy_cls, had_policy, "__main__"), ("Box", box_cls, had_box, "gym.spaces.box")):
            if module in sys.modules:
                if had is None:
                    try:
                        delattr(sys.modules[module], name)
                    except AttributeError:
                        pass
                else:
                    setattr(sys.modules[module], name, had)
        for m in fake:
            sys.modules.pop(m, None)
    if not path.endswith(".pkl"):
        path += ".pkl"                                     # BaseRLModel._save_to_file appends the extension
    with open(path, "wb") as f:
        f.write(blob)
    return path


class FusedLstmPolicy:
    """`act_model` of the reference (n_steps = 1, batch = n_envs; ppo2.py:128-129) on the fused CUDA kernel.
    `step(obs, state, mask)` mirrors CustomLSTMPolicy.step (run_bp_v5.py:178-185): returns
    (actions, values, new_states, neglogps); numpy in -> numpy out, torch CUDA in -> torch CUDA out."""

    def __init__(self, params: Sequence[np.ndarray], n_env: int, device: int = 0, seed: int = 0, env_offset: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.n_env, self.device, self.seed, self.env_offset = int(n_env), int(device), int(seed), int(env_offset)
        flat = flatten_params(params)
        _lib.check(self._L.irrl_policy_create(self.device, C.c_void_p(flat.ctypes.data), C.byref(self._h)), "policy_create")
        self.initial_state = np.zeros((self.n_env, STATE_DIM), dtype=np.float32)     # run_bp_v5.py:174-175
        self.tick = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.irrl_policy_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_params(self, params: Sequence[np.ndarray]) -> None:
        flat = flatten_params(params)
        _lib.check(self._L.irrl_policy_set_params(self._h, C.c_void_p(flat.ctypes.data)), "policy_set_params")

    def set_params_flat_device(self, ptr: int) -> None:
        _lib.check(self._L.irrl_policy_set_params(self._h, C.c_void_p(int(ptr))), "policy_set_params")

    def step(self, obs, state=None, mask=None, deterministic=False, tick: Optional[int] = None, return_clipped=False):
        n = obs.shape[0]
        if tick is None:
            tick = self.tick
            self.tick += 1
        if type(obs).__module__.startswith("torch"):
            import torch
            assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous()
            state = state.clone()
            act = torch.empty((n, 12), device=obs.device); clip = torch.empty((n, 12), device=obs.device)
            val = torch.empty((n,), device=obs.device); nlp = torch.empty((n,), device=obs.device)
            m = None if mask is None else mask.to(torch.uint8).contiguous()
            st = torch.cuda.current_stream(obs.device).cuda_stream
            _lib.check(self._L.irrl_policy_act(self._h, C.c_void_p(st), n, C.c_void_p(obs.data_ptr()), C.c_void_p(m.data_ptr()) if m is not None else None,
                                               C.c_void_p(state.data_ptr()), C.c_void_p(act.data_ptr()), C.c_void_p(clip.data_ptr()), C.c_void_p(val.data_ptr()),
                                               C.c_void_p(nlp.data_ptr()), int(deterministic), self.seed, self.env_offset, int(tick)), "policy_act")
            return (act, val, state, nlp, clip) if return_clipped else (act, val, state, nlp)
        obs = np.ascontiguousarray(obs, np.float32)
        state = np.array(self.initial_state[:n] if state is None else state, dtype=np.float32, order="C", copy=True)
        m = None if mask is None else np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
        act = np.empty((n, 12), np.float32); clip = np.empty((n, 12), np.float32); val = np.empty(n, np.float32); nlp = np.empty(n, np.float32)
        _lib.check(self._L.irrl_policy_act(self._h, None, n, C.c_void_p(obs.ctypes.data), C.c_void_p(m.ctypes.data) if m is not None else None,
                                           C.c_void_p(state.ctypes.data), C.c_void_p(act.ctypes.data), C.c_void_p(clip.ctypes.data), C.c_void_p(val.ctypes.data),
                                           C.c_void_p(nlp.ctypes.data), int(deterministic), self.seed, self.env_offset, int(tick)), "policy_act")
        return (act, val, state, nlp, clip) if return_clipped else (act, val, state, nlp)

    def value(self, obs, state=None, mask=None):
        return self.step(obs, state, mask, deterministic=True, tick=0)[1]
