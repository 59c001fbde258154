"""Pins the oracle's gait generator + IK (Environment.hpp:1756-1890, 1687-1751) against the reference's shipped
trajectory dump Exp_Raw_Data/trot_ref_.csv (fixture tests/golden/trot_ref.npz, made by oracle/gen_golden.py).

Recipe: GaitType 0, period 0.2, lam 0.5, stand_height 0.28, up_height 0.08, WILDCAT False;
t_k = (k+1)*0.002; Vx command through a first-order filter v_k = 0.999 v_{k-1} + 0.001*5.
"""
import os

import numpy as np

from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg
from oracle_lib import Oracle

G = os.path.join(os.path.dirname(__file__), "golden")


def test_trot_ref_known_answer():
    g = np.load(os.path.join(G, "trot_ref.npz"))
    Q, DQ = g["q"], g["dq"]
    o = Oracle(trot_cfg(num_envs=1, num_threads=1, StochasticDynamics=False))
    v = 0.0
    t = 0.0
    worst_q, worst_dq = 0.0, 0.0
    qprev = None
    for k in range(0, 10000):
        v = 0.999 * v + 0.001 * 5.0
        t += 0.002
        # sequential calls like the env: jointRefLast_ is the previous call's jointRef_ (Environment.hpp:1886-1887)
        q, dq, ee = o.gait(t, [v, 0.0, 0.0], is_first=(k == 0))
        worst_q = max(worst_q, np.abs(q - Q[k]).max())
        if k > 0:                   # row 0 of the dump was seeded by an unknown previous sample
            worst_dq = max(worst_dq, np.abs(dq - DQ[k]).max())
    assert worst_q < 1e-4, worst_q                 # CSV carries 6 significant digits
    assert worst_dq < 3e-3, worst_dq               # dq = differences of q's with ~1e-6 rounding / 0.002


def test_joint_dot_ref_is_finite_difference_and_first_step_zero():
    o = Oracle(trot_cfg(num_envs=1, num_threads=1, StochasticDynamics=False))
    q0, dq0, _ = o.gait(0.3, [2.0, 0.0, 0.3], is_first=True)
    q1, dq1, _ = o.gait(0.3, [2.0, 0.0, 0.3], is_first=False)     # same time again: reference quirk 2
    assert np.allclose(q0, q1)
    assert np.abs(dq1).max() == 0.0
    q2, dq2, _ = o.gait(0.302, [2.0, 0.0, 0.3], is_first=False)
    assert np.allclose(dq2, (q2 - q1) / 0.002)
