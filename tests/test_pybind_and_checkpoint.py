"""The compiled pybind11 module (csrc/pybind_shim.cpp: class FlexibleGymEnv as raisim_gym.cpp:14-47 binds it) and the
reference-format checkpoint writer (PPO2.save, ppo2.py:452-476)."""
import ast
import io
import os
import pickle
import re
import sys
import types

import numpy as np
import pytest

from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, dump_yaml
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES, TF_VARIABLE_NAMES, save_reference_pkl, load_reference_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
# the 29 distinct names of raisim_gym.cpp:16-46 (`step` is bound twice there)
PYB_METHODS = ["init", "getExtraInfoNames", "reset", "observe", "step", "setSeed", "testStep", "close", "isTerminalState", "setSimulationTimeStep",
               "setControlTimeStep", "getObDim", "getActionDim", "getExtraInfoDim", "getNumOfEnvs", "startRecordingVideo", "stopRecordingVideo", "showWindow",
               "hideWindow", "curriculumUpdate", "OriginState", "GetOriginStateDim", "ReferenceState", "GetJointEffort", "GetGeneralizedForce",
               "GetInverseMassMatrix", "GetNonlinear", "SetContactCoefficient", "GetSphereInfo"]


def _module():
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import build
    build.build()
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import _flexible_robot_pb
    return _flexible_robot_pb


def test_compiled_module_has_every_method_the_reference_binds():
    m = _module()
    assert all(callable(getattr(m.FlexibleGymEnv, n)) for n in PYB_METHODS)
    if os.path.exists(REF):      # the list above is what the reference's own source says, and what its Python adapter calls
        src = open(os.path.join(REF, "IRRL/FlexibleRobotRaisimGym/flex_gym/env/raisim_gym.cpp")).read()
        assert sorted(set(re.findall(r'\.def\("(\w+)"', src))) == sorted(PYB_METHODS)
        tree = ast.parse(open(os.path.join(REF, "IRRL/FlexibleRobotRaisimGym/flex_gym/env/RaisimGymVecEnv.py")).read())
        used = {n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Attribute) and n.value.attr == "wrapper"}
        missing = [u for u in used if not hasattr(m.FlexibleGymEnv, u) and u != "seed"]      # `wrapper.seed` does not exist in the reference either (RaisimGymVecEnv.py:150)
        assert not missing, missing


def test_compiled_module_refuses_to_work_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = _module()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.FlexibleGymEnv("", dump_yaml(trot_cfg()))


@pytest.mark.gpu
def test_compiled_module_equals_the_ctypes_class_and_rejects_wrong_arrays():
    from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv as CtypesEnv
    from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import RaisimGymVecEnv
    m = _module()
    n = 24
    cfg = dump_yaml(trot_cfg(num_envs=n, StochasticDynamics=True, ObsNoise=2.0))
    a = RaisimGymVecEnv(m.FlexibleGymEnv("", cfg))             # the VecEnv adapter over the compiled class, used like run_bp_v5.py:209 does
    b = RaisimGymVecEnv(CtypesEnv("", cfg))
    assert a.num_obs == 35 and a.num_acts == 12 and a.extra_info_names == b.extra_info_names and a.num_envs == n
    oa, ob = a.reset(), b.reset()
    assert np.array_equal(oa, ob)
    rng = np.random.default_rng(0)
    for t in range(20):
        act = np.clip(rng.normal(0, 0.3, size=(n, 12)), -1, 1).astype(np.float32)
        ra, rb = a.step(act), b.step(act)
        assert all(np.array_equal(x, y) for x, y in zip(ra[:3], rb[:3]))
        assert ra[3][5] == rb[3][5]
    assert np.array_equal(a.OriginState(), b.OriginState()) and np.array_equal(a.GetNonlinear(), b.GetNonlinear())
    assert np.array_equal(a.GetInverseMassMatrix(), b.GetInverseMassMatrix()) and np.array_equal(a.GetJointEffort(), b.GetJointEffort())
    w = a.wrapper
    good = [np.zeros((n, 12), np.float32), np.zeros((n, 35), np.float32), np.zeros(n, np.float32), np.zeros(n, bool), np.zeros((n, 6), np.float32)]
    w.step(*good)
    ro = np.zeros((n, 35), np.float32); ro.setflags(write=False)
    for i, bad in ((0, np.zeros((n, 12), np.float64)), (1, np.zeros((n, 36), np.float32)), (2, np.zeros((n, 1), np.float32)),
                   (1, np.zeros((35, n), np.float32).T), (3, np.zeros(n, np.int32)), (1, ro), (4, [[0.0] * 6] * n)):
        args = list(good); args[i] = bad
        with pytest.raises(TypeError):                                 # Eigen::Ref semantics: no implicit copy, wrong dtype / stride / shape raises
            w.step(*args)
    with pytest.raises(RuntimeError, match="Stiffness"):               # READ_YAML message (RaisimGymEnv.hpp:41-42)
        bad_cfg = trot_cfg(num_envs=2); del bad_cfg["Stiffness"]
        m.FlexibleGymEnv("", dump_yaml(bad_cfg))


def _weights():
    z = np.load(os.path.join(ROOT, "tests", "golden", "bp5_155_params.npz"))
    return [z[k] for k in PARAM_NAMES]


def test_checkpoint_has_the_layout_of_ppo2_save(tmp_path):
    W = [w + 0.01 for w in _weights()]
    path = save_reference_pkl(str(tmp_path / "trained"), W, n_steps=750, learning_rate=5e-5, n_envs=65536)
    assert path.endswith(".pkl")
    back = load_reference_params(path)                                   # the loader that reads the reference's shipped bp5_155.pkl
    assert all(np.array_equal(a, np.asarray(b, np.float32)) for a, b in zip(back, W))

    class Resolve(pickle.Unpickler):                                     # what the reference process provides: the live classes under these names
        def find_class(self, module, name):
            if (module, name) in (("__main__", "CustomLSTMPolicy"), ("gym.spaces.box", "Box")):
                return type(name, (), {})
            return super().find_class(module, name)
    data, params = Resolve(io.BytesIO(open(path, "rb").read())).load()
    assert list(data.keys()) == ["gamma", "n_steps", "vf_coef", "ent_coef", "max_grad_norm", "learning_rate", "lam", "nminibatches", "noptepochs", "cliprange",
                                 "verbose", "policy", "observation_space", "action_space", "n_envs", "_vectorize_action", "policy_kwargs"]      # ppo2.py:453-471
    assert data["n_envs"] == 65536 and data["learning_rate"] == 5e-5 and data["policy_kwargs"] == {"n_lstm": [48, 48]} and data["policy"].__name__ == "CustomLSTMPolicy"
    assert data["observation_space"].shape == (35,) and data["action_space"].shape == (12,) and float(data["action_space"].high[0]) == 1.0
    assert len(params) == 19 and [p.shape for p in params][:3] == [(35, 192), (48, 192), (192,)] and all(p.dtype == np.float32 for p in params)
    if os.path.exists(os.path.join(REF, "IRRL/script/pkl/bp5_155.pkl")):  # same tuple layout and array shapes as the file the reference ships
        ref = load_reference_params(os.path.join(REF, "IRRL/script/pkl/bp5_155.pkl"))
        assert [p.shape for p in ref] == [p.shape for p in params]


@pytest.mark.skipif(not os.path.exists(REF), reason="needs the reference tree (build container)")
def test_exported_checkpoint_round_trips_through_the_references_customer_lstm_nn(tmp_path):
    """CustomerLstmNN(model_path) (CustomerLstmNN.py:27-60) calls PPO2.load(path).get_parameters() and indexes the result by TF
    variable name.  stable-baselines is not installed, so PPO2 is a stand-in that reads the file exactly like BaseRLModel does
    (`data, params = cloudpickle.load(file)`; names = the graph's trainable variables); everything downstream -- weight
    extraction, predict() -- is the reference's unmodified class, compared with the numpy oracle on the same weights."""
    from oracle import lstm_oracle as LO
    rng = np.random.default_rng(7)
    W = [w + rng.normal(0, 0.02, size=w.shape).astype(np.float32) for w in _weights()]
    os.makedirs(tmp_path / "pkl"); os.makedirs(tmp_path / "model")
    save_reference_pkl(str(tmp_path / "pkl" / "mine"), W)

    class PPO2:
        @classmethod
        def load(cls, path):
            class Resolve(pickle.Unpickler):
                def find_class(self, module, name):
                    if (module, name) in (("__main__", "CustomLSTMPolicy"), ("gym.spaces.box", "Box")):
                        return type(name, (), {})
                    return super().find_class(module, name)
            self = cls(); self.data, self.params = Resolve(io.BytesIO(open(path, "rb").read())).load(); return self

        def get_parameters(self):
            return dict(zip(TF_VARIABLE_NAMES, self.params))

    saved = {k: sys.modules.get(k) for k in ("raisim_gym", "raisim_gym.algo", "raisim_gym.algo.ppo2", "utils", "utils.CustomerLstmNN")}
    cwd = os.getcwd()
    try:
        sys.modules["raisim_gym"] = types.ModuleType("raisim_gym"); sys.modules["raisim_gym.algo"] = types.ModuleType("raisim_gym.algo")
        mod = types.ModuleType("raisim_gym.algo.ppo2"); mod.PPO2 = PPO2; sys.modules["raisim_gym.algo.ppo2"] = mod
        sys.modules.pop("utils", None); sys.modules.pop("utils.CustomerLstmNN", None)
        sys.path.insert(0, os.path.join(REF, "IRRL", "script"))
        os.chdir(tmp_path)                                               # the class looks for ./model/<name> and ./pkl/<name>.pkl under the cwd
        from utils.CustomerLstmNN import CustomerLstmNN
        pol = CustomerLstmNN("./pkl/mine.pkl", n_lstm=[48, 48], flag_v=True)
    finally:
        os.chdir(cwd); sys.path.remove(os.path.join(REF, "IRRL", "script"))
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    assert np.array_equal(pol.lstm_wx[0], W[0]) and np.array_equal(pol.v_lstm_wh[1], W[10]) and np.array_equal(pol.pi_w, W[14])
    P = dict(zip(PARAM_NAMES, W)); state = np.zeros((1, 384)); pol.reset()
    for t in range(16):
        x = rng.normal(0, 0.5, size=35)
        got = pol.predict(x)
        a, v, state, nlp, mean = LO.act(P, x[None], state, np.zeros(1))
        assert np.abs(got - np.clip(mean[0], -1, 1)).max() < 1e-6
