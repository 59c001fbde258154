import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
import numpy as np, ctypes as C
from test_gpu_terrain import _pair, N
from oracle_lib import S
from gpu_lib import rel
o,c,h=_pair("perlin")
rng=np.random.default_rng(21)
o.set_tick(1); c.env.setTick(1); o.reset(); c.reset()
for t in range(20):
    s=o.get_state(); c.set_state(s.astype(np.float32))
    a=np.clip(rng.normal(0,0.2,size=(N,12)),-1,1).astype(np.float32)
    obo,ro,do,eo=o.step(a); obg,rg,dg,eg=c.step(a)
    so,sg=o.get_state(),c.get_state()
    err=np.abs(obg-obo).max(axis=1)/np.abs(obo).max()
    cm=(sg[:,S["contact"]]!=so[:,S["contact"]]).any(axis=1)
    idx=np.argsort(-err)[:5]
    print(t, "ncontact", int(so[:,S["contact"]].sum()), "top err", np.round(err[idx],6), "contact mismatch", cm[idx], "argmax col", [int(np.abs(obg[i]-obo[i]).argmax()) for i in idx], "median", np.median(err))
    if t==13:
        i=idx[0]; print("env",i,"gv diff", np.round(sg[i,S["gv"]]-so[i,S["gv"]],6), "contacts", so[i,S["contact"]], "z", so[i,2], "sweeps", o.contact_info(i)["sweeps"], c.sweeps()[i])
