"""GPU bring-up of the tcgen05 act kernel: GEMM probe variants, then tc vs FMA vs numpy-oracle act parity, then timing.
Run each stage in its own process (a trap poisons the CUDA context): python tests/tools/tc_bringup.py probe|act|time"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy, PARAM_NAMES

L = _lib.load()
NOFLUSH = "noflush" in sys.argv
G = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden")


def probe():
    rng = np.random.default_rng(0)
    for (k, n) in [(8, 16), (32, 192), (48, 16), (64, 192)]:
        a = rng.normal(size=(128, k)).astype(np.float32); b = rng.normal(size=(n, k)).astype(np.float32)
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        for variant in (0, 1):
            d = np.zeros((128, n), np.float32)
            rc = L.irrl_tc_gemm_probe(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_void_p(d.ctypes.data), k, n, variant)
            if rc:
                print("probe rc", rc, L.irrl_last_error().decode()); return
            print(f"k={k} n={n} variant={variant}: max|err|={np.abs(d - ref).max():.3e}  (|ref|max {np.abs(ref).max():.2f})", flush=True)


def act():
    from oracle import lstm_oracle as LO
    z = np.load(os.path.join(G, "bp5_155_params.npz")); W = [z[k] for k in PARAM_NAMES]; P = dict(zip(PARAM_NAMES, W))
    for n in (300, 128, 4096):
        rng = np.random.default_rng(n)
        obs = rng.normal(0, 0.7, size=(n, 35)).astype(np.float32); state = rng.normal(0, 0.4, size=(n, 384)).astype(np.float32)
        mask = rng.random(n) < 0.3
        pol = FusedLstmPolicy(W, n_env=n, seed=7, env_offset=3)
        out = {}
        for mode in (1, 2):
            _lib.check(L.irrl_policy_set_act_path(mode))
            out[mode] = pol.step(obs, state, mask, deterministic=False, tick=12, return_clipped=True)
            out[mode + 10] = pol.step(obs, state, mask, deterministic=True, tick=12)
        _lib.check(L.irrl_policy_set_act_path(0))
        names = ["action", "value", "state", "neglogp", "clipped"]
        for i, nm in enumerate(names):
            print(f"n={n} tc-vs-fma {nm}: {np.abs(out[1][i] - out[2][i]).max():.3e}", flush=True)
        ra, rv, rs, rnlp, rmean = LO.act(P, obs, state, mask.astype(np.float64))
        for mode in (11, 12):
            a, v, s, nlp = out[mode]
            print(f"n={n} mode={mode - 10} vs oracle: mean {np.abs(a - rmean).max():.3e} value {np.abs(v - rv).max():.3e} state {np.abs(s - rs).max():.3e} nlp {np.abs(nlp - rnlp).max():.3e}", flush=True)


def timing():
    import torch
    z = np.load(os.path.join(G, "bp5_155_params.npz")); W = [z[k] for k in PARAM_NAMES]
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for n in ((4096, 8192) if NOFLUSH else (4096, 8192, 16384, 32768)):
        pol = FusedLstmPolicy(W, n_env=n)
        obs = torch.randn(n, 35, device=dev) * 0.7; state = torch.randn(n, 384, device=dev) * 0.4
        act = torch.empty(n, 12, device=dev); clip = torch.empty(n, 12, device=dev); val = torch.empty(n, device=dev); nlp = torch.empty(n, device=dev)
        done = torch.zeros(n, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for mode in (1, 2):
            _lib.check(L.irrl_policy_set_act_path(mode))
            ts = []
            for it in range(30):
                if not NOFLUSH: flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(L.irrl_policy_act(pol.handle, C.c_void_p(st), n, C.c_void_p(obs.data_ptr()), C.c_void_p(done.data_ptr()), C.c_void_p(state.data_ptr()),
                                             C.c_void_p(act.data_ptr()), C.c_void_p(clip.data_ptr()), C.c_void_p(val.data_ptr()), C.c_void_p(nlp.data_ptr()), 0, 1, 0, it))
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts = np.array(ts[5:])
            print(f"n={n} mode={'fma' if mode == 1 else 'tc '}: {np.median(ts):.1f} us (min {ts.min():.1f})", flush=True)
            if mode == 2 and n == 4096:
                tl = np.zeros(32, np.int64)
                for rep_ in (1, 1, 1):
                    L.irrl_tc_timeline(rep_, None)
                    if not NOFLUSH: flush.zero_()
                    _lib.check(L.irrl_policy_act(pol.handle, C.c_void_p(st), n, C.c_void_p(obs.data_ptr()), C.c_void_p(done.data_ptr()), C.c_void_p(state.data_ptr()),
                                                 C.c_void_p(act.data_ptr()), C.c_void_p(clip.data_ptr()), C.c_void_p(val.data_ptr()), C.c_void_p(nlp.data_ptr()), 0, 1, 0, 99))
                    L.irrl_tc_timeline(rep_, C.c_void_p(tl.ctypes.data))
                    names = ["start", "A staged", "mma: A0 seen", "mma: W0 seen", "mma: L0 issued", "epi: D0 seen", "epi: L0 done", "mma: A1 seen", "mma: L1 issued",
                             "epi: D1 seen", "epi: L1 done", "mma: A2 seen", "epi: D2 seen", "mma: head issued", "prod: ring filled", "prod: all issued",
                             "pro: obs in", "pro: obs->A", "pro: rows in", "pro: h0->A", "-", "head stored", "stores drained", "entry"]
                    print(f"dbg={rep_} timeline (cycles from start):", ", ".join(f"{nm}={int(tl[i] - tl[0])}" for i, nm in sorted(enumerate(names), key=lambda kv: tl[kv[0]])), flush=True)
                L.irrl_tc_timeline(0, None)


def rate():
    out = np.zeros(2, np.int64)
    for n in (16, 96, 192, 256):
        for name, lt, lbo, sbo, kadv in [("none lbo=2048 sbo=128", 0, 2048, 128, 4096), ("none lbo=128 sbo=256", 0, 128, 256, 0), ("none same-op", 0, 2048, 128, 0),
                                         ("sw32 sbo=256", 6, 0, 256, 4096), ("sw64 sbo=512", 4, 0, 512, 32), ("sw128 sbo=1024", 2, 0, 1024, 32)]:
            for reps in (32, 128):
                _lib.check(L.irrl_tc_mma_rate(n, reps, lt, lbo, sbo, kadv, C.c_void_p(out.ctypes.data)))
                print(f"N={n:3d} {name:24s} reps={reps:3d}: issue {out[0] / reps:6.1f} cyc/MMA, complete {out[1] / reps:6.1f} cyc/MMA", flush=True)


if __name__ == "__main__":
    {"probe": probe, "act": act, "time": timing, "rate": rate}[sys.argv[1]]()
