import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg
from oracle_lib import Oracle, S
from gpu_lib import Cuda, rel
cfg = trot_cfg(num_envs=128, num_threads=8, StochasticDynamics=False, ObsNoise=0.0); cfg.update(Filter=True, Freq=30)
o, c = Oracle(cfg), Cuda(cfg)
rng = np.random.default_rng(0)
o.set_tick(1); c.env.setTick(1)
o.reset(); c.reset()
for t in range(12):
    c.set_state(o.get_state().astype(np.float32))
    a = np.clip(rng.normal(0, 0.25, size=(o.n, 12)), -1, 1).astype(np.float32)
    obo, ro, do, eo = o.step(a); obg, rg, dg, eg = c.step(a)
    so, sg = o.get_state(), c.get_state()
    err = np.abs(obg - obo).max(axis=1) / np.abs(obo).max()
    bad = np.flatnonzero(err > 2e-4)
    sw = np.zeros(o.n, np.int32); c.env.getSolverSweeps(sw)
    print(t, "n_bad", len(bad), "max err", err.max(), "mask diff", int((sg[:, S["contact"]] != so[:, S["contact"]]).any(axis=1).sum()))
    for i in bad[:4]:
        j = np.argmax(np.abs(obg[i] - obo[i]))
        print("   env", i, "err", err[i], "worst ob idx", j, obg[i, j], obo[i, j], "contact", so[i, S["contact"]], sg[i, S["contact"]], "sweeps", sw[i],
              "max|qd|", np.abs(so[i, S["gv"]][6:]).max(), "torque diff", np.abs(sg[i, S["torque"]] - so[i, S["torque"]]).max())
