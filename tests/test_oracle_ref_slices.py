"""CPU: the oracle's restatements against the REFERENCE's own compiled code for every piece of Environment.hpp that builds without
RaiSim / Eigen (oracle/ref_slices_wrap.cpp: sampling_reshape ENV:71-81, gauss ENV:96-99, smooth_function / smooth_function2
ENV:118-156, inverse_kinematics ENV:1687-1751, the torque_clamp arithmetic ENV:1275-1305).

Two checks: against tests/golden/ref_slices.npz (outputs of oracle/_ref/libenv_ref_slices.so, written by
oracle/gen_ref_slices_golden.py; runs everywhere) and, where the built library is present (this container, and the GPU box, to
which oracle/_ref travels), against the library itself."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, test_cfg as manual_test_cfg
from oracle_lib import Oracle, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("gen_ref_slices_golden", os.path.join(ROOT, "oracle", "gen_ref_slices_golden.py"))
G = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(G)


def _oracle_functions():
    L = lib()
    L.bp5o_helper.restype = C.c_double; L.bp5o_helper.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    L.bp5o_ik.argtypes = [C.c_void_p] + [C.c_double] * 3 + [C.c_int, C.c_void_p]
    L.bp5o_torque_clamp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    envs = {}

    def env_for(m):
        key = tuple(m)
        if key not in envs:
            envs[key] = Oracle(trot_cfg(num_envs=1, num_threads=1, MotorMaxTorque=m[0], MotorCriticalSpeed=m[1], MotorMaxSpeed=m[2]))
        return envs[key]

    o0 = env_for((18.0, 100.0, 200.0))

    def ik(x, y, z, right):
        th = np.zeros(3); L.bp5o_ik(o0.h, x, y, z, right, th.ctypes.data); return th

    def clamp(t, g, m):
        t = np.array(t, np.float64); g = np.ascontiguousarray(g, np.float64)
        L.bp5o_torque_clamp(env_for(m).h, t.ctypes.data, g.ctypes.data); return t

    return dict(sampling_reshape=lambda r: L.bp5o_helper(0, r, 0, 0), gauss=lambda x, w, h: L.bp5o_helper(1, x, w, h),
                smooth_function=lambda p, s, l: L.bp5o_helper(2, p, s, l), smooth_function2=lambda p, s, l: L.bp5o_helper(3, p, s, l), ik=ik, clamp=clamp)


def _check(out, ref):
    for k in ("sampling_reshape", "gauss", "smooth_function", "smooth_function2", "torque_clamp"):
        assert out[k].shape == ref[k].shape
        assert np.abs(out[k] - ref[k]).max() <= 1e-15 * max(1.0, np.abs(ref[k]).max()), k       # same double arithmetic: equal to rounding
    # asin / acos near +-1 amplify one ulp of their argument to ~1e-8
    assert np.abs(out["ik"] - ref["ik"]).max() <= 1e-9
    # the fixtures exercise both sides of every branch
    assert (ref["smooth_function"] == 0).any() and (ref["smooth_function"] == 1).any() and ((ref["smooth_function"] > 0) & (ref["smooth_function"] < 1)).any()
    assert (ref["ik"] == 0).any()                                   # infeasible targets: the angle is left untouched ("error1/2/3", ENV:1707-1750)


def test_restated_helpers_match_the_golden_vectors_of_the_reference_code():
    ref = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_slices.npz")))
    assert float(ref["pi"]) == 3.1415926                            # ENV:45
    _check(G.evaluate(_oracle_functions()), ref)


def test_restated_helpers_match_the_compiled_reference_slices():
    if os.path.exists("/root/reference"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(G.SO):
        pytest.skip("oracle/_ref/libenv_ref_slices.so not built (needs /root/reference)")
    L = G.load()
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
    try:
        ref = G.evaluate(G.reference_functions(L))
    finally:
        os.dup2(saved, 1); os.close(devnull)
    _check(G.evaluate(_oracle_functions()), ref)
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_slices.npz")))
    for k in ref:
        assert np.array_equal(ref[k], gold[k]), k                   # the committed fixture is what the library produces
