"""-m gpu: every YAML switch of the hot path (ENV:1616-1629, 1643-1658) against the oracle: action low-pass filter, observation
filter, time-based contact flags, variable foot height, WILDCAT mirroring, gallop phases, action noise, the manual (tele-op test)
configuration of bp5_test.yaml and non-default time steps."""
import numpy as np
import pytest

from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, test_cfg as manual_test_cfg
from oracle_lib import Oracle, S
from gpu_lib import Cuda, rel

pytestmark = pytest.mark.gpu
N = 128


def _run(cfg, steps=12, sigma=0.25, tol=2e-5, seed=0):
    """Teacher-forced single-step comparisons.  Env-steps whose discrete decisions sit within fp32 resolution of a threshold are
    identified from the ORACLE's own margins (oracle_lib.Oracle.margins: a toe / trunk-corner gap within 1e-7 m of zero during the
    substeps, a touch-down speed within 1e-4 m/s of the restitution threshold, an impulse within 0.2 % of the friction cone, the
    termination test within 1e-5 of its bound -- the thresholds measured in tests/test_gpu_same_state.py), set aside and counted
    (<= 3 % of the run); every other env-step must meet the tolerance and ALL flags bit-exactly, with no tolerated exception."""
    o, c = Oracle(cfg), Cuda(cfg)
    rng = np.random.default_rng(seed)
    o.set_tick(1); c.env.setTick(1)
    obo, obg = o.reset(), c.reset()
    assert rel(obg, obo) < max(tol, 5e-5)
    aside = 0
    for t in range(steps):
        c.set_state(o.get_state().astype(np.float32))
        a = np.clip(rng.normal(0, sigma, size=(o.n, 12)), -1, 1).astype(np.float32)
        obo, ro, do, eo = o.step(a); obg, rg, dg, eg = c.step(a)
        so, sg = o.get_state(), c.get_state(); m = o.margins()
        ok = (m[:, 0] > 1e-7) & (m[:, 1] > 1e-4) & (m[:, 2] > 1e-5) & (m[:, 3] > 2e-3)
        aside += int((~ok).sum())
        assert (do[ok] == dg[ok]).all(), t
        assert (sg[ok][:, S["contact"]] == so[ok][:, S["contact"]]).all(), t
        assert rel(obg[ok], obo[ok]) < tol and rel(rg[ok], ro[ok]) < tol and rel(eg[ok], eo[ok]) < tol, (t, rel(obg[ok], obo[ok]), rel(rg[ok], ro[ok]))
        assert rel(sg[ok][:, S["ptarget_last"]], so[ok][:, S["ptarget_last"]]) < tol
        assert rel(sg[ok][:, S["torque_last"]], so[ok][:, S["torque_last"]]) < 1e-4
    assert aside <= max(2, 3 * steps * o.n // 100), aside
    return o, c


def _cfg(**kw):
    d = trot_cfg(num_envs=N, num_threads=8, StochasticDynamics=False, ObsNoise=0.0)
    d.update(kw)
    return d


def test_action_lowpass_filter():            # ENV:396, 703
    _run(_cfg(Filter=True, Freq=30))


def test_observation_filter_is_stateful():   # ENV:1251-1256, 425-426
    cfg = _cfg(ObsFilter=True, ObsNoise=2.0)
    o, c = _run(cfg)
    # observe() applies the filter again on every call (stateful in the reference)
    c.set_state(o.get_state().astype(np.float32))
    a1, b1 = o.observe(), c.observe()
    a2, b2 = o.observe(), c.observe()
    assert rel(b1, a1) < 2e-5 and rel(b2, a2) < 2e-5


def test_time_based_contact_flags():         # ENV:1169-1188
    o, c = _run(_cfg(TimeBasedContact=True))
    assert set(np.unique(c.get_state()[:, S["contact"]])) <= {0.0, 1.0}


def test_height_variable_and_lean():         # ENV:1779-1798
    _run(_cfg(HeightVariable=True, LeanFront=0.01, LeanHind=-0.01, Vy=0.3))


def test_wildcat_bounding_and_gallop():      # ENV:403-409, 589, 1501, 1773
    _run(_cfg(WILDCAT=True, GaitType=1))
    _run(_cfg(GaitType=2, lam=0.4, period=0.25))


def test_action_noise_shared_scalar():       # ENV:704-705
    _run(_cfg(ActionNoise=0.1))


def test_abad_offset_and_gains():            # ENV:317-322, 340-350, 375-380
    _run(_cfg(abad=0.1, AbadRatio=0.7, Stiffness=60.0, Damping=2.0, MotorCriticalSpeed=14.2, MotorMaxSpeed=40))


def test_manual_teleop_configuration():      # bp5_test.yaml: Manual True -> fixed initial state, no command / gait updates
    cfg = manual_test_cfg(num_envs=3, render=False)
    o, c = _run(cfg, steps=10, sigma=0.1)
    sg = c.get_state()
    assert np.abs(sg[:, S["command"]]).max() == 0 and np.abs(sg[:, S["joint_dot_ref"]]).max() == 0


def test_state_disturbance_every_ten_periods():   # ENV:744-747, 912-940 (ForceDisturbance + Manual)
    cfg = manual_test_cfg(num_envs=64, render=False, ForceDisturbance=True)
    o, c = _run(cfg, steps=3, sigma=0.1)
    # frame_idx = 1000 = int(period / control_dt * 10) fires the disturbance: kernel and oracle draw the same Philox numbers
    s0 = o.get_state().copy(); s0[:, S["frame_idx"]] = 1000
    for e in range(o.n):
        o.set_state(e, s0[e])
    c.set_state(s0.astype(np.float32))
    before = s0[:, S["gc"]][:, 3:7].copy()
    a = np.zeros((o.n, 12), np.float32)
    obo, ro, do, eo = o.step(a); obg, rg, dg, eg = c.step(a)
    so, sg = o.get_state(), c.get_state()
    assert np.abs(so[:, S["gc"]][:, 3:7] - before).max() > 1e-3                     # it fired
    ok = (sg[:, S["contact"]] == so[:, S["contact"]]).all(axis=1) & (do == dg)
    assert ok.sum() >= o.n - 1
    assert rel(sg[ok][:, S["gc"]], so[ok][:, S["gc"]]) < 2e-5 and rel(obg[ok], obo[ok]) < 2e-5 and rel(rg[ok], ro[ok]) < 2e-5


def test_meteor_sphere_attack():             # Crutial: True, ENV:717-741, 815-861 (new specification shared with the oracle)
    cfg = _cfg(Crutial=True, CubeNum=2)
    o, c = Oracle(cfg), Cuda(cfg)
    o.set_tick(1); c.env.setTick(1)
    obo, obg = o.reset(), c.reset()
    mo, mg = o.get_meteor(), c.get_meteor()
    assert rel(mg, mo) < 1e-6 and (mo[:, 6] == 0).all() and np.allclose(mo[:, 2], o.get_state()[:, 2] + 1.0)      # placed 1 m above the trunk
    info = np.zeros((o.n, 4), np.float32); c.env.GetSphereInfo(info)
    assert np.allclose(info[:, :3], mo[:, :3], atol=1e-6) and np.allclose(info[:, 3], mo[:, 7], atol=1e-7)
    a = np.zeros((o.n, 12), np.float32)
    hit = hit_checked = 0
    for t in range(140):
        c.set_state(o.get_state().astype(np.float32)); c.set_meteor(o.get_meteor().astype(np.float32))
        vz0 = o.get_meteor()[:, 5].copy(); gv0 = o.get_state()[:, S["gv"]].copy()
        obo, ro, do, eo = o.step(a); obg, rg, dg, eg = c.step(a)
        mo, mg = o.get_meteor(), c.get_meteor()
        so, sg = o.get_state(), c.get_state()
        fresh = do | dg                                   # auto-reset: the oracle re-places the sphere at once, the kernel at its next step
        b_o, b_g = (mo[:, 5] > vz0 + 1.0) & ~fresh, (mg[:, 5] > vz0 + 1.0) & ~fresh      # the sphere's vertical velocity jumped up: it hit
        hit += int(b_o.sum())
        ok = ~fresh & (b_o == b_g)
        assert (b_o != b_g).sum() <= 1, t
        assert rel(mg[ok], mo[ok]) < 2e-5, (t, rel(mg[ok], mo[ok]))                      # the sphere itself: flight, bounce, ground
        # the robot: the synchronized touchdown of 128 identical robots is dense with knife-edge contact decisions (see _run), so
        # the bulk is checked tightly and the tail loosely
        err = np.maximum(np.abs(obg - obo).max(axis=1) / np.abs(obo).max(), np.abs(sg[:, S["gv"]] - so[:, S["gv"]]).max(axis=1) / np.abs(so[:, S["gv"]]).max())
        assert np.median(err[ok]) < 1e-5 and np.percentile(err[ok], 90) < 1e-4, (t, np.median(err[ok]), np.percentile(err[ok], 90))
        # the impulse handed to the robot: where the sphere struck the trunk, the trunk was pushed down by the same amount
        struck = ok & b_o & (mo[:, 2] > so[:, 2])
        if struck.any():
            dvz_o, dvz_g = so[struck][:, S["gv"]][:, 2] - gv0[struck][:, 2], sg[struck][:, S["gv"]][:, 2] - gv0[struck][:, 2]
            good = np.abs(dvz_g - dvz_o) < 2e-3 * np.abs(dvz_o).max()
            assert good.sum() >= struck.sum() - 1 and (dvz_o < 0).mean() > 0.9, (t, dvz_o[:4], dvz_g[:4])
            hit_checked += int(struck.sum())
    assert hit >= o.n // 2 and hit_checked >= o.n // 2     # most spheres did strike their robot within 0.28 s


def test_other_time_steps_loop_count():      # ENV:711
    o, c = _run(_cfg(simulation_dt=0.0005, control_dt=0.004), steps=6)


def test_ragged_counts_with_noise_and_dr():
    for n in (1, 5, 31, 33, 65):
        cfg = trot_cfg(num_envs=n, num_threads=2, StochasticDynamics=True, ObsNoise=2.0)
        _run(cfg, steps=3)
