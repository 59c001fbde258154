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


def test_parity_with_the_oracle_under_a_modified_description(tmp_path):
    """heavier trunk and thighs, longer abad offset, bigger feet: M, h and teacher-forced steps still follow the oracle fed with the same 40 numbers"""
    from gpu_lib import Cuda, rel
    from oracle_lib import Oracle, S
    L = _lib.load()
    builtin = np.zeros(40, np.float32); _lib.check(L.irrl_parse_urdf(None, C.c_void_p(builtin.ctypes.data)))
    m = unpack(builtin.astype(np.float64)); m["m0"] = 5.25; m["m2"] = 0.7; m["off1x"] = 0.23; m["toe_r"] = 0.03; m["com2"] = np.array([0.0, -0.015, -0.03])
    write_urdf(str(tmp_path / "black_panther.urdf"), m)
    got = np.zeros(40, np.float32); _lib.check(L.irrl_parse_urdf(str(tmp_path / "black_panther.urdf").encode(), C.c_void_p(got.ctypes.data)))
    n = 64
    cfg = trot_cfg(num_envs=n, num_threads=4, StochasticDynamics=True, ObsNoise=2.0)
    o = Oracle(cfg, model40=got.astype(np.float64))
    c = Cuda.__new__(Cuda); c.env = FlexibleGymEnv(str(tmp_path), dump_yaml(cfg)); c.env.init(); c.n = n
    o.set_tick(1); c.env.setTick(1)
    obo, obg = o.reset(), c.reset()
    assert rel(obg, obo) < 2e-5
    Mo = np.stack([o.mass_and_h(i)[0] for i in range(4)]); ho = np.stack([o.mass_and_h(i)[1] for i in range(4)])
    assert rel(c.mass_matrix()[:4], Mo) < 1e-5 and rel(c.nonlinear()[:4], ho) < 1e-5
    rng = np.random.default_rng(0); n_out = 0
    for t in range(60):
        c.set_state(o.get_state().astype(np.float32))
        a = np.clip(rng.normal(0, 0.2, size=(n, 12)), -1, 1).astype(np.float32)
        obo, ro, do, eo = o.step(a); obg, rg, dg, eg = c.step(a)
        # same statistical reading as the other contact-rich parity tests: the bulk agrees tightly, envs on the other side of a discrete
        # contact decision (touch, stick / slide, restitution threshold) are outliers bounded in number
        err = np.abs(obg - obo).max(axis=1) / np.abs(obo).max()
        n_out += int((err > 2e-4).sum())
        assert np.median(err) < 1e-5 and np.percentile(err, 85) < 1e-4, (t, np.median(err), np.percentile(err, 85))
        assert (do != dg).sum() <= 2, t
    assert n_out <= 0.05 * 60 * n
