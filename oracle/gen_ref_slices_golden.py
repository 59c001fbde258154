#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Writes tests/golden/ref_slices.npz: outputs of the REFERENCE's own code for the pure-arithmetic pieces of
Environment.hpp that compile without RaiSim / Eigen (oracle/Makefile target `ref` -> oracle/_ref/libenv_ref_slices.so; see
oracle/ref_slices_wrap.cpp for the line ranges).  Run in the build container, where /root/reference exists:

    make -C oracle ref && python oracle/gen_ref_slices_golden.py
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libenv_ref_slices.so")


def load():
    L = C.CDLL(SO)
    for n, k in (("ref_sampling_reshape", 1), ("ref_gauss", 3), ("ref_smooth_function", 3), ("ref_smooth_function2", 3), ("ref_sgn", 1)):
        f = getattr(L, n); f.restype = C.c_double; f.argtypes = [C.c_double] * k
    L.ref_pi.restype = C.c_double
    L.ref_inverse_kinematics.argtypes = [C.c_double] * 7 + [C.c_void_p, C.c_int]
    L.ref_torque_clamp.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    return L


def inputs():
    """the fixed input sets shared by the generator and the test"""
    rng = np.random.default_rng(2024)
    ratio = np.concatenate([np.linspace(0.0, 1.0, 401), [0.5, 0.4999999, 0.5000001]])
    phase = np.concatenate([np.linspace(0.0, 3.0, 1201), rng.uniform(0, 50, 400)])
    lam = np.array([0.3, 0.5, 0.7]); slope = np.array([2.0, 0.4])
    gx = np.linspace(-0.2, 1.2, 281)
    # IK targets: reachable swing/stance toe positions, targets beyond max_len, targets that make asin/acos arguments leave [-1, 1]
    ik = np.concatenate([np.column_stack([rng.uniform(-0.25, 0.25, 600), rng.uniform(-0.15, 0.15, 600), rng.uniform(-0.42, -0.12, 600)]),
                         np.column_stack([rng.uniform(-0.6, 0.6, 100), rng.uniform(-0.4, 0.4, 100), rng.uniform(-0.6, -0.01, 100)])])
    right = (np.arange(len(ik)) % 2).astype(np.int32)
    tq = rng.uniform(-40, 40, size=(300, 12)); gv = np.zeros((300, 18)); gv[:, 6:] = rng.uniform(-260, 260, size=(300, 12))
    motors = np.array([[18.0, 100.0, 200.0], [18.0, 14.2, 40.0]])          # default_cfg.yaml:35-37, bp5_test.yaml
    return dict(ratio=ratio, phase=phase, lam=lam, slope=slope, gx=gx, ik=ik, ik_right=right, tq=tq, gv=gv, motors=motors)


def evaluate(fn):
    """fn: dict of callables with the reference's signatures -> dict of outputs on inputs()"""
    I = inputs(); out = {}
    out["sampling_reshape"] = np.array([fn["sampling_reshape"](r) for r in I["ratio"]])
    out["gauss"] = np.array([[fn["gauss"](x, 1.0, h) for x in I["gx"]] for h in (0.08, 0.05)])
    for name in ("smooth_function", "smooth_function2"):
        out[name] = np.array([[[fn[name](p, s, l) for p in I["phase"]] for l in I["lam"]] for s in I["slope"]])
    th = np.zeros((len(I["ik"]), 3))
    for i, (p, r) in enumerate(zip(I["ik"], I["ik_right"])):
        th[i] = fn["ik"](p[0], p[1], p[2], int(r))
    out["ik"] = th
    out["torque_clamp"] = np.array([[fn["clamp"](t, g, m) for t, g in zip(I["tq"], I["gv"])] for m in I["motors"]])
    return out


def reference_functions(L):
    l_hip, l_thigh, l_calf = 0.085, 0.209, 0.2175                           # ENV:1949-1952
    max_len = float(np.sqrt(l_hip * l_hip + (l_calf + l_thigh) ** 2))       # ENV:395

    def ik(x, y, z, right):
        th = np.zeros(3); L.ref_inverse_kinematics(x, y, z, l_hip, l_thigh, l_calf, max_len, th.ctypes.data, right); return th

    def clamp(t, g, m):
        t = np.array(t, np.float64); g = np.ascontiguousarray(g, np.float64); up = np.zeros(18); lo = np.zeros(18)
        L.ref_torque_clamp(t.ctypes.data, g.ctypes.data, m[0], m[1], m[2], up.ctypes.data, lo.ctypes.data); return t

    return dict(sampling_reshape=L.ref_sampling_reshape, gauss=L.ref_gauss, smooth_function=L.ref_smooth_function,
                smooth_function2=L.ref_smooth_function2, ik=ik, clamp=clamp)


if __name__ == "__main__":
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    L = load()
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)     # the reference prints "error1/2/3" on infeasible IK targets
    try:
        out = evaluate(reference_functions(L))
    finally:
        os.dup2(saved, 1)
    out["pi"] = np.array(L.ref_pi())
    path = os.path.join(ROOT, "tests", "golden", "ref_slices.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
