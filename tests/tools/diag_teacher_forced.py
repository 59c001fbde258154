"""diagnostic (GPU box): teacher-forced step comparison, per-env error against the oracle's decision margins"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, relaxation_cfg
from oracle_lib import Oracle, S
from gpu_lib import Cuda
N = 512
cfg = (relaxation_cfg if (len(sys.argv) > 2 and sys.argv[2] == "relaxation") else trot_cfg)(num_envs=N, num_threads=8, StochasticDynamics=True, ObsNoise=2.0)
o, c = Oracle(cfg), Cuda(cfg)
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 5); o.set_tick(1); c.env.setTick(1); o.reset(); c.reset()
rows = []
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 80):
    c.set_state(o.get_state().astype(np.float32))
    a = np.clip(rng.normal(0, 0.3, size=(N, 12)), -1, 1).astype(np.float32)
    obo, ro, do, eo = o.step(a); obg, rg, dg, eg = c.step(a)
    so, sg = o.get_state(), c.get_state(); m = o.margins()
    err = (np.abs(obg.astype(np.float64) - obo) / np.maximum(np.abs(obo), 1.0)).max(axis=1)
    flag = (do != dg) | (sg[:, S["contact"]] != so[:, S["contact"]]).any(axis=1)
    sw = c.sweeps()
    for i in range(N):
        rows.append((t, i, err[i], flag[i], m[i, 0], m[i, 1], m[i, 2], m[i, 3], sw[i], int(so[i, S["contact"]].sum())))
R = np.array(rows)
print("env-steps", len(R), "flag mismatches", int(R[:, 3].sum()), "err>1e-4", int((R[:, 2] > 1e-4).sum()), "err>2e-5", int((R[:, 2] > 2e-5).sum()))
for name, col, ths in (("geo", 4, (1e-7, 3e-7, 1e-6, 2e-6)), ("rest", 5, (1e-5, 1e-4, 1e-3)), ("term", 6, (1e-5,)), ("cone", 7, (1e-5, 1e-4, 1e-3, 1e-2))):
    for th in ths:
        print(f"  share with {name} margin < {th:g}: {np.mean(R[:, col] < th):.4f}")
bad = R[(R[:, 2] > 2e-5) | (R[:, 3] > 0)]
print("t env err flag geo rest term cone sweeps ncontact")
for r in bad[np.argsort(-bad[:, 2])][:60]:
    print("%3d %4d %.1e %d %.1e %.1e %.1e %.1e %2d %d" % tuple(r))
ok = R[(R[:, 4] > 1e-7) & (R[:, 5] > 1e-4) & (R[:, 6] > 1e-5)]
print("worst safe rows (geo>1e-7, no cone filter):")
for r in ok[np.argsort(-ok[:, 2])][:8]:
    print("%3d %4d %.1e %d %.1e %.1e %.1e %.1e %2d %d" % tuple(r))
print("safe share", len(ok) / len(R), "worst err among safe", ok[:, 2].max(), "flags among safe", int(ok[:, 3].sum()))
print("percentiles of err among safe: p50 %.1e p99 %.1e p99.9 %.1e" % tuple(np.percentile(ok[:, 2], [50, 99, 99.9])))
