#!/usr/bin/env python
"""bench.py -- env-steps/s of the bp5 hot path (LSTM act + vectorised env step) on N B200s, and the PPO iteration wall-time.

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement of the reference's OpenMP path on the host cores

Headline workload (BASELINE.json configs[2]): bp5 relaxation phase (mimic reward removed), 16384 environments per B200 with the
fused LSTM act rollout.  One "step" = one pass of the hot path over the environments of one GPU: the LSTM act kernel
(CustomLSTMPolicy.step, run_bp_v5.py:178-185) followed by the env step kernel (VectorizedEnvironment::step,
VectorizedEnvironment.hpp:268-278: PD + 8 physics substeps + observation + reward + termination + auto-reset) with the rollout stores
of Runner.run (ppo2.py:519-538).  Environments shard over GPUs with no data-path collective (weak scaling: 16384 per GPU at every N).

Sub-records in the same JSON line: `configs` -- configs[1] (4096 envs, step kernel + obs/reward only), configs[4]-like stairs +
domain randomisation, and at 8 GPUs configs[3] (8192 per GPU = 65536 envs, the north-star size); `ppo_iteration` -- one full PPO
iteration at 8192 envs per GPU (750-step rollout, GAE, 10 BPTT epochs, NCCL gradient all-reduce when N > 1; seconds, and the device
time of the all-reduce).  `e2e` = the same metric through the C ABI with HOST buffers; `roofline` (FP32 CUDA-core bound, SURVEY.md 8d;
HBM fraction beside it); `cpu_baseline` (the oracle's OpenMP step + numpy LSTM act on the host cores, threads stated); `clocks`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG", "WARN")              # keep NCCL's version banner off stdout (rank 0 prints ONE JSON line) unless the caller asks for more
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")      # CPU arm: OpenMP workers must not spin while numpy's BLAS threads run the LSTM act

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "env-steps/sec incl. LSTM act"
UNIT = "env-steps/s"
# Executed FP32 work per env-step of the step kernel, frozen from the ncu capture in profiles/ (see DESIGN.md "Roofline accounting"):
COUNTS_FILE = os.path.join(ROOT, "profiles", "kernel_counts.json")
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 74.5: 148 SMs x 128 FMA lanes x 2 x max SM clock
B_STEP_BYTES = 1.0e3                                     # ~0.77 KB state + DR/counters per env-step
B_ACT_BYTES = 3.3e3                                      # obs 140 + LSTM state 2x1536 + outputs 60
WORKLOAD_TEXT = {"trot": "bp5 trot imitation reward (BASELINE.json configs[1])",
                 "relaxation": "bp5 relaxation phase, mimic reward removed (BASELINE.json configs[2])",
                 "stairs": "bp5 stair heightfield terrain + mass/friction domain randomisation (BASELINE.json configs[4])"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=16384, help="BASELINE.json configs[2]: 16384 envs on 1xB200")
    ap.add_argument("--workload", default="relaxation", choices=["trot", "relaxation", "stairs"],
                    help="configs[2] relaxation (default) | configs[1] trot imitation | configs[4] stair heightfield + DR")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[1] / stairs / 65536-env sub-records")
    ap.add_argument("--no-ppo", action="store_true", help="skip the PPO iteration record")
    ap.add_argument("--ppo-envs", type=int, default=8192, help="envs per GPU of the PPO iteration record (configs[3]: 65536 / 8)")
    ap.add_argument("--ppo-steps", type=int, default=750)
    ap.add_argument("--ppo-epochs", type=int, default=10)
    ap.add_argument("--cpu-envs", type=int, default=2048, help="bounded sample of the workload for the CPU arm")
    return ap.parse_args()


def workload_cfg(workload, n_envs, threads=None):
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, relaxation_cfg
    # flat ground (or stairs), ObsNoise 2.0, domain randomisation on, command / clock / pose from reset
    kw = dict(num_envs=n_envs, num_threads=threads or os.cpu_count() or 1, StochasticDynamics=True, ObsNoise=2.0)
    if workload == "relaxation":      # configs[2]: JointRewardCoeff = EndEffectorRewardCoeff = 0
        return relaxation_cfg(**kw)
    if workload == "stairs":          # configs[4]: stair heightfield (rise 0.08 / run 0.3 m) + mass/friction randomisation
        return trot_cfg(Terrain=True, terrain_kind="stairs", stair_rise=0.08, stair_run=0.3, **kw)
    return trot_cfg(**kw)


def policy_weights():
    from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES, init_params
    g = os.path.join(ROOT, "tests", "golden", "bp5_155_params.npz")
    if os.path.exists(g):
        z = np.load(g)
        return [z[k] for k in PARAM_NAMES], "bp5_155 (reference checkpoint)"
    return init_params(np.random.default_rng(0)), "random-init"


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """NVML is initialised in the constructor (before the warm-up); the thread samples every 2 ms while `active`, and one sample is
    taken synchronously at both ends of the timed region so that even a few-millisecond region is covered."""
    def __init__(self, device_index):
        self.samples, self.reasons, self._stop, self.active = [], set(), threading.Event(), False
        self.dev = device_index; self.max_mhz = None; self.h = None; self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(device_index))
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.names = {pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                          pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        except Exception:
            self.h = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def sample(self):
        if self.h is None:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.dev}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.samples.append(int(out[0])); self.max_mhz = int(out[1])
            except Exception:
                pass
            return
        self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, n in self.names.items():
            if r & bit:
                self.reasons.add(n)

    def _run(self):
        while not self._stop.is_set():
            if self.active and self.h is not None:
                try:
                    self.sample()
                except Exception:
                    pass
            time.sleep(0.002)

    def begin(self):
        self.sample(); self.active = True

    def end(self):
        self.active = False; self.sample()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}

    def close(self):
        self._stop.set(); self.thread.join(timeout=2)


# ----------------------------------------------------------------------------------------------- CPU arm (oracle)
class CpuArm:
    """The reference's OpenMP vec-env step restated (oracle/, built with the reference's own flags -O3 -mtune=native; RaiSim is not
    available) + the LSTM act (CustomerLstmNN.py:112-156 restated as an fp32 OpenMP loop), on the host cores, both phases on the same thread count."""
    def __init__(self, workload, n_envs, threads):
        from oracle_lib import Oracle
        from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES
        self.n, self.threads = n_envs, threads
        self.o = Oracle(workload_cfg(workload, n_envs, threads), fast=True)
        from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import flatten_params
        W, self.wname = policy_weights(); self.flat = flatten_params(W)
        self.obs = self.o.reset(); self.state = np.zeros((n_envs, 384), np.float32); self.done8 = np.zeros(n_envs, np.uint8)
        self.act = np.zeros((n_envs, 12), np.float32); self.val = np.zeros(n_envs, np.float32); self.nlp = np.zeros(n_envs, np.float32)
        self.o.L.bp5o_lstm_act.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_int]; self.o.L.bp5o_lstm_act.restype = None
        self.rng = np.random.default_rng(0)

    def step(self):
        # LSTM act: fp32 C++ / OpenMP loop over the envs (oracle_capi.cpp bp5o_lstm_act, checked against the numpy restatement of
        # CustomerLstmNN.py in tests/test_oracle_lstm_kat.py) -- the multi-threaded stand-in for the reference's TF session.run
        n = self.n
        eps = self.rng.standard_normal((n, 12), dtype=np.float32)
        self.o.L.bp5o_lstm_act(self.flat.ctypes.data_as(C.c_void_p), n, self.obs.ctypes.data_as(C.c_void_p), self.state.ctypes.data_as(C.c_void_p),
                               self.done8.ctypes.data_as(C.c_void_p), eps.ctypes.data_as(C.c_void_p), self.act.ctypes.data_as(C.c_void_p),
                               self.val.ctypes.data_as(C.c_void_p), self.nlp.ctypes.data_as(C.c_void_p), self.threads)
        self.obs, rew, done, extra = self.o.step(np.clip(self.act, -1, 1))
        self.done8 = done.astype(np.uint8)

    def run(self, steps=None, budget_s=None, warmup=2):
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=self.threads):
            for _ in range(warmup):
                self.step()
            t0 = time.perf_counter(); k = 0
            while True:
                self.step(); k += 1
                if (steps is not None and k >= steps) or (budget_s is not None and time.perf_counter() - t0 > budget_s and k >= 3):
                    break
            dt = time.perf_counter() - t0
        return self.n * k / dt, k, dt


def cpu_thread_candidates():
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return cores, sorted({cores, max(1, cores // 2)}, reverse=True)


def best_cpu_arm(workload, n_envs, budget_s):
    """sweep the thread count (all hardware threads / half of them: SMT siblings rarely help this FP-bound loop), keep the best"""
    cores, cands = cpu_thread_candidates()
    best = None; tried = {}
    for th in cands:
        arm = CpuArm(workload, n_envs, th)
        v, k, dt = arm.run(budget_s=budget_s / len(cands))
        tried[str(th)] = v
        if best is None or v > best[0]:
            best = (v, th, k, arm)
    return best, cores, tried


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.cpu_envs
    (v0, th, k0, arm), cores, tried = best_cpu_arm(args.workload, n, budget_s=6.0)
    K = max(1, min(args.steps, 200)); Wm = min(max(args.warmup, 1), 20)
    val, K, total = arm.run(steps=K, warmup=Wm)
    sample = (f"{K} control steps x {n} envs of the {args.workload} workload; oracle restatement of the reference's OpenMP step (RaiSim unavailable), "
              f"fp64 like the reference, built -O3 -mtune=native like CMakeLists.txt:28, + OpenMP fp32 LSTM act; {th} threads of {cores} hardware threads (sweep {tried})")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload] + f" with LSTM act: bounded sample of {n} envs on the host cores", "envs": n, "policy": arm.wname},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "threads": th, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
class Rig:
    """one env shard + policy + device rollout buffers of one workload on this rank's GPU"""
    def __init__(self, L, workload, N, K, rank, local, stream, dev, with_act=True):
        import torch
        from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
        from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import dump_yaml
        from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy
        self.L, self.N, self.K, self.dev, self.stream, self.with_act, self.torch = L, N, K, dev, stream, with_act, torch
        self.env = FlexibleGymEnv("", dump_yaml(workload_cfg(workload, N)), device=local, env_offset=rank * N)
        self.env.setStream(stream.cuda_stream)
        self.env.init()
        W, self.wname = policy_weights()
        self.pol = FusedLstmPolicy(W, n_env=N, device=local, seed=1, env_offset=rank * N)
        f32 = dict(device=dev, dtype=torch.float32)
        self.buf_obs = torch.empty((K, N, 35), **f32); self.buf_act = torch.empty((K, N, 12), **f32); self.buf_val = torch.empty((K, N), **f32)
        self.buf_nlp = torch.empty((K, N), **f32); self.buf_rew = torch.empty((K, N), **f32); self.buf_done = torch.empty((K, N), device=dev, dtype=torch.uint8)
        self.cur_obs = torch.zeros((N, 35), **f32); self.cur_done = torch.zeros((N,), device=dev, dtype=torch.uint8); self.state = torch.zeros((N, 384), **f32)
        self.env.reset(self.cur_obs)
        if not with_act:   # configs[1]: "step kernel + obs/reward only": actions ~ clip(N(0, 0.1^2)) (the trained log-std), resident in HBM
            g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
            self.fixed_act = torch.clamp(torch.randn((8, N, 12), generator=g, **f32) * 0.1, -1, 1)
            self.buf_extra = torch.empty((N, 6), **f32)

    def one(self, t):
        from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
        if self.with_act:
            b = _lib.RolloutBuffers(obs=self.buf_obs[t].data_ptr(), actions=self.buf_act[t].data_ptr(), values=self.buf_val[t].data_ptr(), neglogps=self.buf_nlp[t].data_ptr(),
                                    rewards=self.buf_rew[t].data_ptr(), dones=self.buf_done[t].data_ptr(), cur_obs=self.cur_obs.data_ptr(), cur_done=self.cur_done.data_ptr(),
                                    state=self.state.data_ptr(), ep_return=None, ep_length=None)
            _lib.check(self.L.irrl_rollout(self.env.handle, self.pol.handle, 1, C.byref(b), 0), "rollout")
        else:
            self.env.step(self.fixed_act[t % 8], self.cur_obs, self.buf_rew[t], self.buf_done[t], self.buf_extra)

    def timed(self, Wm, flush, dist, sampler=None):
        """W warm-up steps, then exactly K steps between CUDA events on the launching stream, L2 flushed between steps.
        Returns per-rank total ms (sum over steps), per-kernel ms (act, step) and the wall time."""
        torch = self.torch; K = self.K
        for t in range(Wm):
            self.one(t % K)
        torch.cuda.synchronize(self.dev)
        self.L.irrl_set_profiling(self.env.handle, 1)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(self.dev)
        if sampler is not None:
            sampler.begin()
        wall0 = time.perf_counter()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for t in range(K):
            if flush is not None:
                flush.fill_(t & 0xFF)
            ev[t][0].record(self.stream)
            self.one(t)
            ev[t][1].record(self.stream)
        torch.cuda.synchronize(self.dev)
        if dist is not None:
            dist.barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.end() if sampler is not None else None
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        act_ms, step_ms, cnt = C.c_double(), C.c_double(), C.c_int64()
        self.L.irrl_get_profile(self.env.handle, C.byref(act_ms), C.byref(step_ms), C.byref(cnt))
        self.L.irrl_set_profiling(self.env.handle, 0)
        if not self.with_act and cnt.value == 0:
            act_ms.value, step_ms.value, cnt.value = 0.0, total_ms, K
        t_ms = torch.tensor([total_ms], device=self.dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return dict(total_ms_max=float(t_ms.item()), total_ms=total_ms, act_ms=act_ms.value / max(cnt.value, 1), step_ms=step_ms.value / max(cnt.value, 1), wall=wall, clocks=clocks)


def run_b200(args):
    import torch
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local); torch.cuda.set_device(dev)
    sampler = ClockSampler(local)                      # NVML up before any warm-up
    L = _lib.load()
    N = args.envs_per_gpu; K = args.steps; Wm = max(args.warmup, 3)
    # one explicit (non-default) stream carries everything: env/policy kernels, the L2 flush and the timing events
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    flush = None if args.no_l2_flush else torch.empty(256 << 20, device=dev, dtype=torch.uint8)

    rig = Rig(L, args.workload, N, K, rank, local, stream, dev)
    m = rig.timed(Wm, flush, dist, sampler)
    value = world * N * K / (m["total_ms_max"] * 1e-3)
    sw = np.zeros(N, np.int32); rig.env.getSolverSweeps(sw)
    episodes_done = int(rig.buf_done.sum().item()); mean_rew = float(rig.buf_rew.mean().item())

    # ---- e2e: the reference-facing calls with HOST buffers (model.step + env.step), copies inside the timed region
    Ke = max(10, min(args.e2e_steps, K))
    # RaisimGymVecEnv / the policy object own their host buffers as one page-locked block each, laid out like the native output
    # blocks (vec_env.py, _lib.pinned_block), so every call returns its results in a single DMA copy
    h_obs, h_rew, h_extra, h_done = _lib.pinned_block(((N, 35), np.float32), ((N,), np.float32), ((N, 6), np.float32), ((N,), np.bool_))
    h_act, h_clip, h_val, h_nlp = _lib.pinned_block(((N, 12), np.float32), ((N, 12), np.float32), ((N,), np.float32), ((N,), np.float32))
    env, pol, state = rig.env, rig.pol, rig.state
    env.reset(h_obs)
    state.zero_()
    fused_entry = hasattr(L, "irrl_act_step")
    act_args = (C.c_void_p(stream.cuda_stream), N, C.c_void_p(h_obs.ctypes.data), C.c_void_p(h_done.view(np.uint8).ctypes.data), C.c_void_p(state.data_ptr()),
                C.c_void_p(h_act.ctypes.data), C.c_void_p(h_clip.ctypes.data), C.c_void_p(h_val.ctypes.data), C.c_void_p(h_nlp.ctypes.data))

    def e2e_step(t):
        # model.step(obs, states, dones): obs/mask host -> device, actions/values/neglogp device -> host; LSTM state stays
        # an opaque device-resident handle (the Runner only threads `states` through, ppo2.py:520)
        _lib.check(L.irrl_policy_act(pol.handle, *act_args, 0, 1, rank * N, 100000 + t), "policy_act")
        env.step(h_clip, h_obs, h_rew, h_done, h_extra)     # RaisimGymVecEnv.step -> wrapper.step (RaisimGymVecEnv.py:31)

    def timed_host_loop(fn):
        for t in range(3):
            fn(t)
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for t in range(Ke):
            fn(3 + t)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        return world * N * Ke / float(t_e.item())

    e2e_two_calls = timed_host_loop(e2e_step)
    e2e_value, e2e_path = e2e_two_calls, "irrl_policy_act + irrl_step (the reference's two calls per step: model.step, env.step)"
    h2d = N * (35 * 4 + 1 + 12 * 4)                        # obs + mask (act)  + clipped action (step)
    d2h = N * (12 * 4 * 2 + 4 + 4 + 35 * 4 + 4 + 1 + 6 * 4)  # action, clipped, value, neglogp | obs, reward, done, extra
    e2e_extra = {}
    if fused_entry:
        # one host entry for act -> clip -> step: a single H2D (obs + mask) and a single D2H block per control step
        from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import fused_host_step
        fs = fused_host_step(env, pol, state, stream.cuda_stream, seed=1, env_offset=rank * N)
        env.reset(fs.obs); state.zero_()
        e2e_fused = timed_host_loop(lambda t: fs(100000 + t))
        e2e_extra = {"two_call_value": e2e_two_calls, "fused_value": e2e_fused}
        if e2e_fused > e2e_value:
            e2e_value, e2e_path = e2e_fused, "irrl_act_step: one host entry per control step (act -> clip -> step), one H2D + one D2H of page-locked blocks"
            h2d, d2h = fs.h2d_bytes, fs.d2h_bytes

    # ---- sub-records on BASELINE's other configs (short runs, same timing method)
    subs = {}
    Ks = max(20, min(K, 200))
    try:
        counts0 = json.load(open(COUNTS_FILE))
    except Exception:
        counts0 = {}
    if not args.no_extras:
        def sub(workload, n, with_act, label):
            r = Rig(L, workload, n, Ks, rank, local, stream, dev, with_act=with_act)
            mm = r.timed(Wm, flush, dist)
            subs[label] = {"workload": WORKLOAD_TEXT[workload] + (" with fused LSTM act" if with_act else ", step kernel + obs/reward only (actions resident in HBM)"),
                           "envs_per_gpu": n, "total_envs": world * n, "steps": Ks, "value": world * n * Ks / (mm["total_ms_max"] * 1e-3), "unit": UNIT,
                           "ms_per_step": mm["total_ms_max"] / Ks, "step_kernel_ms": mm["step_ms"], "act_kernel_ms": mm["act_ms"] if with_act else None}
            if not with_act:   # one Python call per step sits between the timing events; the kernel itself is timed by events around its launch
                subs[label]["value_kernel_only"] = n * world / (mm["step_ms"] * 1e-3)
                subs[label]["fp32_frac"] = (counts0.get("env_step", {}).get("fp32_flop_per_env_step", 0) * n / (mm["step_ms"] * 1e-3) / 1e12) / NOMINAL_FP32_TFLOPS
            del r
        if world == 1:
            sub("trot", 4096, False, "configs[1]")
        sub("stairs", {1: 16384, 2: 16384, 4: 8192, 8: 4096}.get(world, 4096), True, "configs[4]")      # 32768 envs at 2/4/8 GPUs
        if world == 8:
            sub("relaxation", 8192, True, "configs[3]-rollout")                                          # 65536 envs, the north-star size

    # ---- PPO iteration wall-time (BASELINE.json metric, second half; configs[3] per-GPU size)
    ppo = None
    if not args.no_ppo:
        ppo = ppo_iteration(args, rank, local, world, dist, dev)

    if rank != 0:
        sampler.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0
    # ---- roofline of the dominant kernel (env_step): FP32 CUDA-core bound, HBM fraction beside it
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"
    fp32_meas = C.c_double(0.0)
    L.irrl_measure_fp32_peak(local, C.byref(fp32_meas))
    counts = {}
    try:
        counts = json.load(open(COUNTS_FILE))
    except Exception:
        pass
    step_kernel_ms, act_kernel_ms = m["step_ms"], m["act_ms"]
    flop_per_env_step = counts.get("env_step", {}).get("fp32_flop_per_env_step")
    per_env = counts.get("env_step", {}).get("dram_bytes_per_env_step")
    traffic = int(per_env * N) if per_env else counts.get("env_step", {}).get("dram_bytes_per_launch")      # ncu dram__bytes of one launch, scaled to this batch
    roof = {"kernel": "env_step_kernel", "bound": "fp32", "unit": "TFLOP/s", "peak": fp32_meas.value or NOMINAL_FP32_TFLOPS,
            "peak_source": "measured live: register-resident FMA micro-benchmark (irrl_measure_fp32_peak); nominal 74.5 (MEASURED_PEAKS.json holds no FP32-SIMT peak)",
            "kernel_ms": step_kernel_ms, "share_of_step": step_kernel_ms / (m["total_ms"] / K), "traffic": traffic,
            "traffic_note": counts.get("env_step", {}).get("dram_note"),
            "hbm": {"achieved": N * B_STEP_BYTES / (step_kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src,
                    "frac": N * B_STEP_BYTES / (step_kernel_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_env_step": B_STEP_BYTES}}
    if flop_per_env_step:
        ach = N * flop_per_env_step / (step_kernel_ms * 1e-3) / 1e12
        roof["oracle_flop_per_env_step"] = counts.get("oracle", {}).get("flop_per_env_step")   # dense CPU formulation, reported only
        roof.update({"achieved": ach, "frac": ach / roof["peak"], "frac_meaning": "FP32-pipe utilisation: EXECUTED thread-level FP32 operations of the kernel (ncu opcode counters) / peak",
                     "flop_per_env_step": flop_per_env_step, "flop_source": counts.get("env_step", {}).get("source", "ncu")})
    else:
        roof.update({"achieved": None, "frac": None, "flop_per_env_step": None, "flop_source": "profiles/kernel_counts.json missing"})
    roof["lstm_act"] = {"kernel_ms": act_kernel_ms, "bound": "hbm", "achieved": N * B_ACT_BYTES / (act_kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": N * B_ACT_BYTES / (act_kernel_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_env": B_ACT_BYTES,
                        "kernel": "lstm_act_tc_kernel: tcgen05.mma kind::tf32 x3 split, TMEM accumulators, cp.async.bulk operand/state staging" if N >= 256 else "lstm_act_kernel (fp32 FMA)",
                        "tensor_tflops_issued": N * 2 * 3 * 2 * (88 * 192 + 96 * 192) / (act_kernel_ms * 1e-3) / 1e12 if N >= 256 else None}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": m["total_ms_max"] / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload] + f", {N} envs per B200"
                                   + " with fused LSTM act rollout: act kernel + env step kernel (PD, 8 physics substeps, obs, reward, done, auto-reset) + rollout stores",
                       "envs_per_gpu": N, "total_envs": world * N, "substeps": 8, "policy": rig.wname, "obs_noise": 2.0, "domain_randomisation": True,
                       "l2": "flushed between timed steps (256 MiB write)" if flush is not None else "not flushed (state << L2)",
                       "parallelism": f"env-shard x{world}, no data-path collective"},
            "e2e": dict({"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                         "path": e2e_path + "; page-locked host numpy buffers (as RaisimGymVecEnv owns them); LSTM state device-resident"}, **e2e_extra),
            "gpu_launches": 2 * K, "kernels": ["lstm_act_tc_kernel (tcgen05 3xTF32)" if N >= 256 else "lstm_act_kernel", "env_step_kernel"], "memcpy_d2d_per_step": 0,
            "roofline": roof, "clocks": m["clocks"], "wall_s": m["wall"], "configs": subs, "ppo_iteration": ppo,
            "sanity": {"mean_reward": mean_rew, "episodes_finished": episodes_done, "gs_sweeps_last_substep": {"mean": float(sw.mean()), "max": int(sw.max()), "hist": np.bincount(sw, minlength=9).tolist()}}}
    sampler.close()
    if not args.no_cpu_baseline and world == 1:
        (v, th, k, arm), cores, tried = best_cpu_arm(args.workload, args.cpu_envs, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "threads": th, "kind": "port",
                                "sample": f"{k} control steps x {args.cpu_envs} envs of the same workload (OpenMP fp32 LSTM act + oracle OpenMP step, fp64 like the reference, -O3 -mtune=native like CMakeLists.txt:28); "
                                          f"thread sweep {tried}"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def ppo_iteration(args, rank, local, world, dist, dev):
    """one full PPO iteration (ppo2.py:363-447): T-step rollout on the device, GAE, noptepochs BPTT epochs over the whole batch with the
    gradient all-reduce over NCCL (N > 1).  Two iterations run, the second is reported."""
    import torch
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import ppo2 as P2
    from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import dump_yaml
    from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import RaisimGymVecEnv
    n, T, E = args.ppo_envs, args.ppo_steps, args.ppo_epochs
    torch.cuda.empty_cache()
    W, wname = policy_weights()
    env = RaisimGymVecEnv(FlexibleGymEnv("", dump_yaml(workload_cfg("relaxation", n)), device=local, env_offset=rank * n))
    model = P2.PPO2(env, policy_params=W, n_steps=T, noptepochs=E, learning_rate=1e-4, verbose=0)
    P2.ALLREDUCE_EVENTS = [] if world > 1 else None
    hist = model.learn(total_timesteps=n * T * world * 2)
    torch.cuda.synchronize(dev)
    h = hist[-1]
    ar = None
    if P2.ALLREDUCE_EVENTS:
        ev = P2.ALLREDUCE_EVENTS[len(P2.ALLREDUCE_EVENTS) // 2:]          # the second iteration's calls
        us = [1e3 * a.elapsed_time(b) for a, b in ev]
        ar = {"calls_per_iteration": len(ev), "bytes": 70741 * 4, "mean_us": float(np.mean(us)), "max_us": float(np.max(us)), "total_ms_per_iteration": float(np.sum(us)) / 1e3}
    P2.ALLREDUCE_EVENTS = None
    t = torch.tensor([h["iteration_s"], h["rollout_s"], h["update_s"]], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    it, ro, up = (float(x) for x in t.tolist())
    del model, env
    torch.cuda.empty_cache()
    return {"metric": "PPO iteration wall-time", "unit": "s", "value": it, "rollout_s": ro, "update_s": up, "envs_per_gpu": n, "total_envs": n * world, "n_steps": T,
            "noptepochs": E, "nminibatches": 1, "samples_per_iteration": n * world * T, "env_steps_per_s_incl_learning": n * world * T / it,
            "gradient_allreduce": ar, "workload": "bp5 relaxation, PPO2 (clipped surrogate, BPTT over 750 steps), weights bp5_155",
            "learner_kernels": "sequence-persistent BPTT recurrence on warp-level tensor-core MMAs (fp16 2-term / tf32 3-term splits = fp32-grade results), streaming projections and weight gradients on tcgen05 + TMEM (3xTF32), heads + PPO loss + gradients in one kernel", "timing": "host wall clock around rollout + update with a device synchronise, max over ranks"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
