"""Configuration of the bp5 environment: the flat ``environment:`` map of the reference.

Every key read by the reference's ``parameter_load_from_yaml`` (Environment.hpp:1594-1659) and by
``VectorizedEnvironment::init`` (VectorizedEnvironment.hpp:136-171) is mandatory here as well; keys the
reference's C++ ignores (Spring_Stiff, Spring_Damp, AbadOffset, Lean_middle, FrontRearOffset) are accepted
and ignored.  ``train_cfg()`` carries the values of the shipped ``urdf/default_cfg.yaml`` (default_cfg.yaml:4-62)
and ``test_cfg()`` those of ``script/config/bp5_test.yaml`` (bp5_test.yaml:4-64); ``trot_cfg()`` is BASELINE.json's
"bp5 trot imitation" configuration: the training file with GaitType 0 / WILDCAT False and the imitation reward
coefficients of bp5_test.yaml:44-51.
"""
from __future__ import annotations

import copy
from typing import Any, Dict

import yaml

#: keys that parameter_load_from_yaml / VectorizedEnvironment::init abort on when missing
REQUIRED_KEYS = (
    "num_envs", "num_threads", "simulation_dt", "control_dt", "seedd",
    "abad", "period", "lam", "stand_height", "up_height", "down_height", "gait_step", "Vx", "Vy", "Omega",
    "LeanFront", "LeanHind", "Terrain", "Manual", "Crutial", "Filter", "Camera", "StochasticDynamics",
    "HeightVariable", "TimeBasedContact", "ManualTraj", "MotorDynamics", "ObsFilter", "WILDCAT",
    "ForceDisturbance", "Convert2Torque", "terminalRewardCoeff", "EndEffectorRewardCoeff", "BodyPosRewardCoeff",
    "BodyAttitudeRewardCoeff", "JointRewardCoeff", "VelRewardCoeff", "TorqueCoeff", "ContactCoeff", "Stiffness",
    "Stiffness_Low", "AbadRatio", "Damping", "Freq", "max_time", "CubeNum", "FPS", "ActionNoise", "ObsNoise",
    "GaitType", "MotorMaxTorque", "MotorCriticalSpeed", "MotorMaxSpeed",
)

_TRAIN: Dict[str, Any] = dict(
    seedd=1, render=False, num_envs=200, num_threads=120, simulation_dt=0.00025, control_dt=0.002, max_time=1.5,
    abad=0.0, period=0.2, lam=0.5, stand_height=0.28, up_height=0.08, down_height=0.0, gait_step=0.15,
    Manual=False, Terrain=False, Filter=False, Crutial=False, Camera=False, StochasticDynamics=True,
    HeightVariable=False, TimeBasedContact=False, ManualTraj=True, MotorDynamics=False, ObsFilter=False,
    WILDCAT=True, ForceDisturbance=False, Convert2Torque=False, GaitType=1, Freq=30,
    MotorMaxTorque=18.0, MotorCriticalSpeed=100, MotorMaxSpeed=200, AbadRatio=1.0, Stiffness=40.0,
    Stiffness_Low=40.0, Damping=1.0, Spring_Stiff=10000, Spring_Damp=0.1, terminalRewardCoeff=-1.0,
    EndEffectorRewardCoeff=0.0, BodyPosRewardCoeff=0.05, BodyAttitudeRewardCoeff=0.05, JointRewardCoeff=0.1,
    VelRewardCoeff=0.6, TorqueCoeff=0.3, ContactCoeff=0.0, Vx=5.0, Vy=0.0, Omega=1.0, AbadOffset=0.0,
    LeanFront=0.0, LeanHind=-0.0, ActionNoise=0.0, ObsNoise=2.0, CubeNum=6, FPS=60.0, RefTraj="",
)

_TEST_OVERRIDES: Dict[str, Any] = dict(
    seedd=10, num_envs=1, num_threads=1, stand_height=0.30, Manual=True, StochasticDynamics=False,
    HeightVariable=True, WILDCAT=False, GaitType=0, MotorCriticalSpeed=14.2, MotorMaxSpeed=40,
    Spring_Stiff=2000, Spring_Damp=0.5, terminalRewardCoeff=-0.0, BodyPosRewardCoeff=0.2,
    BodyAttitudeRewardCoeff=0.2, JointRewardCoeff=0.4, VelRewardCoeff=0.2, TorqueCoeff=0.1, ContactCoeff=0.1,
    Lean_middle=0.0, FrontRearOffset=-0.05, ObsNoise=0.0, CubeNum=1, FPS=100.0,
)

_IMITATION_COEFFS = dict(BodyPosRewardCoeff=0.2, BodyAttitudeRewardCoeff=0.2, JointRewardCoeff=0.4,
                         VelRewardCoeff=0.2, TorqueCoeff=0.1, ContactCoeff=0.1)


def train_cfg(**overrides: Any) -> Dict[str, Any]:
    """``environment:`` map of the shipped training file (bounding, WILDCAT) with render off."""
    d = copy.deepcopy(_TRAIN)
    d.update(overrides)
    return d


def test_cfg(**overrides: Any) -> Dict[str, Any]:
    """``environment:`` map of script/config/bp5_test.yaml (manual single-env tele-op configuration)."""
    d = copy.deepcopy(_TRAIN)
    d.update(_TEST_OVERRIDES)
    d.update(overrides)
    return d


def trot_cfg(**overrides: Any) -> Dict[str, Any]:
    """BASELINE.json "bp5 trot imitation": training file + GaitType 0, WILDCAT False, imitation coefficients."""
    d = train_cfg(GaitType=0, WILDCAT=False, **_IMITATION_COEFFS)
    d.update(overrides)
    return d


def relaxation_cfg(**overrides: Any) -> Dict[str, Any]:
    """BASELINE.json "relaxation phase": the mimic terms removed (readme.md:71-78)."""
    d = trot_cfg(JointRewardCoeff=0.0, EndEffectorRewardCoeff=0.0)
    d.update(overrides)
    return d


def dump_yaml(env_map: Dict[str, Any]) -> str:
    """What run_bp_v5.py:205-207 passes to the native constructor: the dumped ``environment`` sub-map."""
    return yaml.safe_dump(dict(env_map), default_flow_style=False, sort_keys=False)


def load_yaml_file(path: str) -> Dict[str, Any]:
    with open(path, "r") as f:
        doc = yaml.safe_load(f)
    return doc


def to_kv_string(env_map: Dict[str, Any]) -> str:
    """Numeric ``key=value;`` form (booleans as 0/1, strings dropped) used by test infrastructure."""
    items = []
    for k, v in env_map.items():
        if isinstance(v, bool):
            items.append(f"{k}={int(v)}")
        elif isinstance(v, (int, float)):
            items.append(f"{k}={float(v)!r}")
    return ";".join(items)
