"""-m gpu: FlexibleGymEnv(resourceDir, cfg) reads <resourceDir>/black_panther.urdf like the reference (ENV:231) and the kernels use it."""
import ctypes as C

import numpy as np
import pytest

from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, dump_yaml
from urdf_template import unpack, write_urdf

pytestmark = pytest.mark.gpu


def test_resource_dir_description_reaches_the_kernels(tmp_path):
    L = _lib.load()
    builtin = np.zeros(40, np.float32); _lib.check(L.irrl_parse_urdf(None, C.c_void_p(builtin.ctypes.data)))
    m = unpack(builtin.astype(np.float64)); m["m0"] = 5.25; m["m2"] = 0.7
    write_urdf(str(tmp_path / "black_panther.urdf"), m)
    n = 4
    cfg = dump_yaml(trot_cfg(num_envs=n, StochasticDynamics=False, ObsNoise=0.0))
    plain, heavy = FlexibleGymEnv("", cfg), FlexibleGymEnv(str(tmp_path), cfg)
    for e in (plain, heavy):
        e.init()
    mp, mh = np.zeros((n, 94), np.float32), np.zeros((n, 94), np.float32)
    plain.getModelParams(mp); heavy.getModelParams(mh)
    assert np.allclose(mp[:, 3], 3.72) and np.allclose(mh[:, 3], 5.25)                          # trunk mass as seen by the kernels
    Mp, Mh = np.zeros((n, 324), np.float32), np.zeros((n, 324), np.float32)
    plain.GetMassMatrix(Mp); heavy.GetMassMatrix(Mh)
    total_p = 3.72 + 4 * (m["m1"] + 0.636 + m["m3"]); total_h = 5.25 + 4 * (m["m1"] + 0.7 + m["m3"])
    assert np.allclose(Mp[:, 0], total_p, rtol=1e-5) and np.allclose(Mh[:, 0], total_h, rtol=1e-5)   # M[0,0] = total mass
    # a description the compact model cannot represent is refused at construction
    write_urdf(str(tmp_path / "black_panther.urdf"), unpack(builtin.astype(np.float64)), tweak=lambda leg, d: d.update(m1=d["m1"] * 1.5) if leg == "hl" else None)
    with pytest.raises(RuntimeError, match="mirror"):
        FlexibleGymEnv(str(tmp_path), cfg)
