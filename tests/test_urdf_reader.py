"""CPU: the robot description reader behind FlexibleGymEnv(resourceDir, ...) (ENV:231; csrc/urdf_reader.h through irrl_parse_urdf, no GPU
needed).  The built-in constants must be exactly what the reader extracts from a description with those numbers; edits must come through;
a description the compact (mirror-symmetric) model cannot represent must be refused."""
import ctypes as C
import os

import numpy as np
import pytest

from urdf_template import unpack, write_urdf

REF = "/root/reference/IRRL/FlexibleRobotRaisimGym/flex_gym/env/env/BlackPanther_V55/urdf/black_panther.urdf"


def _lib():
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import build, _lib
    build.build()
    return _lib.load()


def _parse(L, path):
    out = np.zeros(40, np.float32)
    rc = L.irrl_parse_urdf(path.encode() if path else None, C.c_void_p(out.ctypes.data))
    return rc, out, L.irrl_last_error().decode()


def test_round_trip_of_the_builtin_model(tmp_path):
    L = _lib()
    rc, builtin, _ = _parse(L, None)
    b = unpack(builtin)
    assert rc == 0 and b["m0"] == np.float32(3.72) and b["toe_r"] == np.float32(0.0275)          # trunk mass, toe radius (URDF:18, 148)
    p = str(tmp_path / "black_panther.urdf")
    write_urdf(p, unpack(builtin.astype(np.float64)))
    rc, got, err = _parse(L, p)
    assert rc == 0, err
    assert np.abs(got - builtin).max() <= 1e-6 * np.abs(builtin).max()


def test_edits_come_through_and_asymmetry_is_refused(tmp_path):
    L = _lib()
    _, builtin, _ = _parse(L, None)
    m = unpack(builtin.astype(np.float64)); m["m0"] = 5.25; m["box_half"] = np.array([0.2, 0.12, 0.06]); m["rotor"] = np.array([0.004, 0.004, 0.009])
    p = str(tmp_path / "black_panther.urdf")
    write_urdf(p, m)
    rc, got, err = _parse(L, p)
    g = unpack(got.astype(np.float64))
    assert rc == 0 and abs(g["m0"] - 5.25) < 1e-6 and np.allclose(g["box_half"], [0.2, 0.12, 0.06]) and np.allclose(g["rotor"], [0.004, 0.004, 0.009])
    # a heavier right-front thigh: the four legs are no longer mirror images
    write_urdf(p, unpack(builtin.astype(np.float64)), tweak=lambda leg, d: d.update(m2=d["m2"] * 1.2) if leg == "fr" else None)
    rc, _, err = _parse(L, p)
    assert rc != 0 and "mirror" in err
    rc, _, err = _parse(L, str(tmp_path / "missing.urdf"))
    assert rc != 0 and "cannot open" in err


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree is only mounted in the build container")
def test_the_shipped_description_gives_the_builtin_constants():
    """pins the hard-coded model of the kernels (capi.cu model_defaults) on the reference's own black_panther.urdf, incl. its mirror symmetry"""
    L = _lib()
    _, builtin, _ = _parse(L, None)
    rc, got, err = _parse(L, REF)
    assert rc == 0, err
    assert np.array_equal(got, builtin)


def test_oracle_accepts_the_same_compact_model():
    """the CPU oracle can be built on the 40 numbers of a description (parity tests under a modified robot, tests/test_gpu_urdf.py)"""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from oracle_lib import Oracle
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg
    L = _lib()
    _, builtin, _ = _parse(L, None)
    cfg = trot_cfg(num_envs=2, num_threads=1, StochasticDynamics=False, ObsNoise=0.0)
    base, same = Oracle(cfg), Oracle(cfg, model40=builtin.astype(np.float64))
    m = builtin.astype(np.float64).copy(); m[25] = 5.25          # trunk mass
    heavy = Oracle(cfg, model40=m)
    for o in (base, same, heavy):
        o.set_tick(1); o.reset()
    M0, Ms, Mh = base.mass_and_h(0)[0], same.mass_and_h(0)[0], heavy.mass_and_h(0)[0]
    assert abs(M0[0, 0] - (3.72 + 4 * (0.54 + 0.636 + 0.114))) < 1e-9 and np.abs(Ms - M0).max() < 1e-6        # float32 rounding of the 40 numbers
    assert abs(Mh[0, 0] - M0[0, 0] - (5.25 - 3.72)) < 1e-6
