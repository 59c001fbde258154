"""CPU: what the contact solver of the oracle (and of the CUDA kernel, which follows the same specification) converges to.

The RaiSim boundary is unpinned (closed source), so this file pins the solver against RaiSim's *published* algorithm instead
(Hwangbo, Lee, Hutter, RA-L 2018; call site ENV:768, material ENV:433, ERP 0 ENV:246): per-contact Gauss-Seidel sweeps to
convergence, each visit solving the single-contact problem exactly on the friction cone (bisection on the cone boundary).

* schedule: the product runs the foot contacts of a sweep simultaneously (block Jacobi) for the first 10 sweeps because that maps
  onto one lane per leg, then one after the other like the trunk-box corners; <= 30 sweeps, relative tolerance 1e-5.
  `solver_jacobi=0, solver_iters=500, solver_tol=1e-12` is plain per-contact Gauss-Seidel to convergence.  On > 10^4
  contact-rich substeps the two give the same post-impact velocity to <= 1e-5 (belly-down poses with 6-8 contacts: 1e-4) and
  no sample reaches the sweep cap -- under either single-contact rule.  (Round 1 ran pure block Jacobi capped at 10 sweeps:
  `jacobi_sweeps=99, solver_iters=10` below shows why that was changed -- it does not converge for 0.03 % of the bounding
  substeps with 3-4 coupled feet and stops early for 16 % of the belly-down poses.)
* single-contact rule: the product updates the sliding direction with ONE fixed-point step from the stick direction
  (`slide_iters=1`); `slide_exact=1` (oracle only) finds the maximal-dissipation direction by bisection like RaiSim.  That is
  an approximation, not a converged solve: this file measures it per substep (percentiles below) and at the level the
  north-star states for contact-rich behaviour (rollout statistics within 2 %).
"""
import os

import numpy as np
import pytest

from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, train_cfg, test_cfg as manual_test_cfg
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES
from oracle_lib import Oracle, S
from oracle import lstm_oracle as LO
from gpu_lib import stance_states

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GS_CONVERGED = dict(solver_jacobi=0, solver_iters=500, solver_tol=1e-12)


def _policy():
    z = np.load(os.path.join(ROOT, "tests", "golden", "bp5_155_params.npz"))
    return {k: z[k] for k in PARAM_NAMES}


def _contact_rich_states(n=128, steps=110, skip=30):
    """(state, joint torque) samples: the reference's trained policy trotting (touch-downs, stance, sliding at lift-off), the
    shipped bounding configuration (robots that stumble, fall and lie on their trunk), upright stance and belly-down poses."""
    P = _policy(); rng = np.random.default_rng(0)
    out_s, out_t = [], []
    for cfg in (trot_cfg(num_envs=n, num_threads=8, StochasticDynamics=True, ObsNoise=0.0),
                train_cfg(num_envs=n, num_threads=8, ObsNoise=0.0)):
        o = Oracle(cfg); o.set_tick(1)
        obs = o.reset(); state = np.zeros((n, 384)); done = np.zeros(n, bool)
        for t in range(steps):
            a, v, state, nlp, mean = LO.act(P, obs, state, done.astype(np.float64), rng.normal(size=(n, 12)))
            obs, r, done, _ = o.step(np.clip(a, -1, 1).astype(np.float32))
            if t >= skip:
                st = o.get_state(); out_s.append(st); out_t.append(st[:, 179:191].copy())
    st = stance_states(rng, 256); out_s.append(st); out_t.append(rng.uniform(-10, 10, size=(256, 12)))
    st = stance_states(rng, 256)                                   # trunk-box contacts: belly down, legs folded, rolled
    st[:, 2] = rng.uniform(0.03, 0.09, size=256); st[:, 7:19] = np.tile([0.0, -1.4, 2.6], 4)
    roll = rng.uniform(-0.4, 0.4, size=256); st[:, 3] = np.cos(roll / 2); st[:, 4] = np.sin(roll / 2); st[:, 5:7] = 0
    st[:, 19:37] = rng.normal(size=(256, 18)) * 0.3
    out_s.append(st); out_t.append(np.zeros((256, 12)))
    cat = np.concatenate([np.full(len(x), min(i // (steps - skip), 2) if i < 2 * (steps - skip) else (2 if i == 2 * (steps - skip) else 3)) for i, x in enumerate(out_s)])
    return np.concatenate(out_s), np.concatenate(out_t), cat   # cat: 0 trot rollouts, 1 bounding rollouts, 2 stance, 3 belly-down


@pytest.fixture(scope="module")
def solver_samples():
    """every variant integrates ONE substep from the same injected state with the same torques"""
    states, taus, cats = _contact_rich_states()
    base = dict(num_envs=1, num_threads=1, StochasticDynamics=False, ObsNoise=0.0)
    variants = dict(product=dict(), gs=dict(GS_CONVERGED), product_exact=dict(slide_exact=1), gs_exact=dict(slide_exact=1, **GS_CONVERGED),
                    round1=dict(jacobi_sweeps=99, solver_iters=10))
    O = {k: Oracle(trot_cfg(**base, **v)) for k, v in variants.items()}
    rows = {k: [] for k in O}
    rows["cat"] = []
    for s, tau, cat in zip(states, taus, cats):
        res = {}
        for k, o in O.items():
            o.set_state(0, s); o.integrate(0, tau)
            res[k] = (o.get_state(0)[S["gv"]].copy(), o.contact_info(0))
        if res["gs_exact"][1]["n_contacts"] == 0:
            continue
        rows["cat"].append(cat)
        for k in O:
            rows[k].append((res[k][0], res[k][1]["sweeps"], res[k][1]["n_contacts"], res[k][1]["foot_impulse"].copy()))
    return rows


def _rel_du(rows, a, b):
    return np.array([np.abs(x[0] - y[0]).max() / max(np.abs(y[0]).max(), 1e-9) for x, y in zip(rows[a], rows[b])])


def test_block_jacobi_schedule_equals_converged_per_contact_gauss_seidel(solver_samples):
    rows = solver_samples
    n = len(rows["product"])
    assert n >= 10000, n                                               # >= 10^4 substeps with at least one contact
    assert np.mean([r[2] for r in rows["product"]]) > 1.5              # and mostly several contacts at once
    for jac, gs in (("product", "gs"), ("product_exact", "gs_exact")):
        du = _rel_du(rows, jac, gs)
        dl = np.array([np.abs(x[3] - y[3]).max() / max(np.abs(y[3]).max(), 1e-6) for x, y in zip(rows[jac], rows[gs])])
        sweeps = np.array([r[1] for r in rows[jac]])
        print(f"{jac} vs {gs}: n={n} du+ p50 {np.median(du):.1e} p99 {np.percentile(du, 99):.1e} max {du.max():.1e}; impulses max {dl.max():.1e}; "
              f"sweeps mean {sweeps.mean():.2f} max {sweeps.max()} at cap {np.mean(sweeps >= 30):.4f}")
        belly = np.array(rows["cat"]) == 3
        assert du[~belly].max() <= 1e-5 and dl[~belly].max() <= 1e-5   # the GPU tolerance of the step comparison
        # 6-8 contacts on one rigid body: the 1e-5 stopping rule leaves 2e-5 (one bisection-rule sample stops at the cap with 7e-4)
        assert du[belly].max() <= (1e-4 if jac == "product" else 1e-3)
        if jac == "product":
            assert np.mean(sweeps >= 30) == 0.0                        # the sweep cap never binds
    old = _rel_du(rows, "round1", "gs")
    print(f"round-1 schedule (pure block Jacobi, 10 sweeps): max {old.max():.1e}, share > 1e-5 {np.mean(old > 1e-5):.4f}")
    assert old.max() > 1e-3                                            # the defect this round removed is real


def test_one_step_sliding_direction_is_an_approximation_of_the_exact_cone_solve(solver_samples):
    """measured, not hidden: how far the product's single-contact rule is from RaiSim's exact rule, per substep"""
    du = _rel_du(solver_samples, "product", "product_exact")
    p50, p90, p99 = np.percentile(du, [50, 90, 99])
    print(f"one-step sliding rule vs bisection: du+ p50 {p50:.1e} p90 {p90:.1e} p99 {p99:.1e} max {du.max():.1e}; share > 1e-5: {np.mean(du > 1e-5):.3f}")
    assert p50 < 1e-7            # sticking / separating contacts: identical
    assert p90 < 5e-3 and p99 < 0.1
    assert np.mean(du > 1e-5) < 0.45


def _anchor(seconds=4.0, n=4, vx_cmd=5.0, **kw):
    """scripts/eval_bp5_155.py on the oracle: bp5_155 driven like run_bp_v5.py --test, command ramped towards vx_cmd"""
    P = _policy()
    o = Oracle(manual_test_cfg(num_envs=n, num_threads=4, render=False, stand_height=0.28, friction=0.8, **kw))
    obs = o.reset(); state = np.zeros((n, 384)); done = np.zeros(n, bool)
    T = int(seconds / 0.002); v = 0.0; rec = np.zeros((T, n, 37)); falls = 0
    for t in range(T):
        v = 0.999 * v + 0.001 * vx_cmd
        obs[:, 0] = v - 2.5; obs[:, 1] = 0; obs[:, 2] = 0
        a, val, state, nlp, mean = LO.act(P, obs, state, done.astype(np.float64), np.zeros((n, 12)))
        obs, r, done, _ = o.step(np.clip(mean, -1, 1).astype(np.float32)); falls += int(done.sum())
        rec[t] = o.get_state()[:, :37]
    h = rec[T // 2:]
    sig = h[:, 0, 8] - h[:, 0, 8].mean(); f = np.fft.rfftfreq(len(sig), 0.002); stride = float(f[1:][np.argmax(np.abs(np.fft.rfft(sig))[1:])])
    return dict(falls=falls, vx=float(h[:, :, 19].mean()), z=float(h[:, :, 2].mean()), stride_hz=stride)


def test_rollout_statistics_product_rule_vs_exact_rule_within_2_percent():
    """north_star: contact-rich rollouts are checked statistically (within 2 %).  Reference RaiSim rollout of the same policy:
    z 0.2732 m, stride 5 Hz (Exp_Raw_Data/body-center-2021-06-22-16-48-33.bin)."""
    a = _anchor()
    b = _anchor(slide_exact=1, **GS_CONVERGED)
    print("product rule", a, "| exact rule, converged Gauss-Seidel", b)
    assert a["falls"] == 0 and b["falls"] == 0
    assert abs(a["vx"] - b["vx"]) < 0.02 * abs(b["vx"])
    assert abs(a["z"] - b["z"]) < 0.02 * b["z"]
    assert abs(a["stride_hz"] - b["stride_hz"]) <= 0.26
    assert abs(a["z"] - 0.2732) < 0.02 * 0.2732 and abs(b["z"] - 0.2732) < 0.02 * 0.2732
