import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg
from oracle_lib import Oracle, S
from gpu_lib import Cuda, rel, stance_states
N = 1024
rng = np.random.default_rng(2)
s = stance_states(rng, N); tau = rng.uniform(-10, 10, size=(N, 12))
cfg0 = trot_cfg(num_envs=N, num_threads=8, StochasticDynamics=False, ObsNoise=0.0)
o = Oracle(dict(cfg0, solver_tol=1e-9, solver_iters=50))
for i in range(N): o.set_state(i, s[i]); o.integrate(i, tau[i])
ref = o.get_state(); osw = np.array([o.contact_info(i)["sweeps"] for i in range(N)])
oimp = np.stack([o.contact_info(i)["foot_impulse"] for i in range(N)])
print("oracle(1e-9) sweeps hist", np.bincount(osw))
for tol in (1e-6, 1e-5, 3e-5, 1e-4, 1e-3):
    c = Cuda(dict(cfg0, solver_tol=tol)); c.set_state(s.astype(np.float32))
    act, imp = c.integrate(tau); g = c.get_state(); sw = c.sweeps()
    scale = np.abs(oimp).max(axis=(1, 2)) + 1e-6
    ierr = np.abs(imp - oimp).max(axis=(1, 2)) / scale
    verr = np.abs(g[:, S["gv"]] - ref[:, S["gv"]]).max(axis=1) / np.abs(ref[:, S["gv"]]).max(axis=1)
    print(f"tol {tol:g}: sweeps hist {np.bincount(sw)} mean {sw.mean():.2f} | impulse relerr median {np.median(ierr):.2e} max {ierr.max():.2e} | gv relerr max {verr.max():.2e}")
for tol in (1e-7, 1e-6, 1e-5, 1e-4):
    o2 = Oracle(dict(cfg0, solver_tol=tol))
    for i in range(N): o2.set_state(i, s[i]); o2.integrate(i, tau[i])
    sw = np.array([o2.contact_info(i)["sweeps"] for i in range(N)])
    imp = np.stack([o2.contact_info(i)["foot_impulse"] for i in range(N)])
    ierr = np.abs(imp - oimp).max(axis=(1, 2)) / (np.abs(oimp).max(axis=(1, 2)) + 1e-6)
    print(f"oracle tol {tol:g}: sweeps mean {sw.mean():.2f} hist {np.bincount(sw)} impulse relerr max {ierr.max():.2e}")
