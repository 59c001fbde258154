#!/usr/bin/env python
"""bench.py -- env-steps/s of the bp5 hot path (LSTM act + vectorised env step) on N B200s.

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement of the reference's OpenMP path on the host cores

One "step" = one pass of the hot path over the batch of environments of one GPU: the fused LSTM act kernel
(CustomLSTMPolicy.step, run_bp_v5.py:178-185) followed by the env step kernel (VectorizedEnvironment::step,
VectorizedEnvironment.hpp:268-278: PD + 8 physics substeps + observation + reward + termination + auto-reset), with the
rollout stores of Runner.run (ppo2.py:519-538).  Environments shard over GPUs with no data-path collective (weak scaling).

Printed JSON (one line, rank 0): metric/value/unit/..., `e2e` (host buffers through the C ABI, host<->device copies
inside the timed region), `roofline` (FP32 CUDA-core bound; HBM fraction reported beside it),
`cpu_baseline` (the oracle's OpenMP step on the host cores), `clocks`, `gpu_launches`.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "env-steps/sec incl. LSTM act"
UNIT = "env-steps/s"
# Algorithmic work per env-step, frozen from the ncu capture in profiles/ (see DESIGN.md "Roofline accounting"):
COUNTS_FILE = os.path.join(ROOT, "profiles", "kernel_counts.json")
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 74.5: 148 SMs x 128 FMA lanes x 2 x max SM clock
B_STEP_BYTES = 1.0e3                                     # ~0.77 KB state + DR/counters per env-step
B_ACT_BYTES = 3.3e3                                      # obs 140 + LSTM state 2x1536 + outputs 60


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=4096, help="BASELINE.json configs[1]: 4096 envs on 1xB200")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--cpu-envs", type=int, default=200, help="reference CPU configuration (default_cfg.yaml:7)")
    ap.add_argument("--workload", default="trot", choices=["trot", "relaxation", "stairs"],
                    help="BASELINE.json configs[1] trot imitation (default) | configs[2] relaxation (mimic reward removed) | configs[4] stair heightfield + DR")
    return ap.parse_args()


WORKLOAD = "trot"


def workload_cfg(n_envs):
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, relaxation_cfg
    # BASELINE.json configs[1] / SURVEY.md 8d config 2: trot imitation coefficients, flat ground, ObsNoise 2.0, DR on
    kw = dict(num_envs=n_envs, num_threads=os.cpu_count() or 1, StochasticDynamics=True, ObsNoise=2.0)
    if WORKLOAD == "relaxation":      # configs[2]: JointRewardCoeff = EndEffectorRewardCoeff = 0
        return relaxation_cfg(**kw)
    if WORKLOAD == "stairs":          # configs[4]: stair heightfield (rise 0.08 / run 0.3 m) + mass/friction randomisation
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
    def __init__(self, device_index):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.dev = device_index
        self.max_mhz = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.dev)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.002)
        except Exception as e:   # NVML missing: fall back to one nvidia-smi query
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.dev}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.samples.append(int(out[0])); self.max_mhz = int(out[1])
            except Exception:
                pass

    def start(self):
        self.thread.start()

    def stop(self):
        self._stop.set(); self.thread.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_arm(n_envs, budget_s, with_act=True, seed=0):
    """The reference's OpenMP vec-env step restated (oracle/, RaiSim is not available) + numpy LSTM act, timed on the
    host cores.  Returns (env-steps/s, cores, description)."""
    from oracle_lib import Oracle
    from oracle import lstm_oracle as LO
    from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES
    cores = os.cpu_count() or 1
    cfg = workload_cfg(n_envs); cfg["num_threads"] = cores
    o = Oracle(cfg)
    W, _ = policy_weights(); P = dict(zip(PARAM_NAMES, W))
    obs = o.reset(); state = np.zeros((n_envs, 384)); done = np.zeros(n_envs, bool)
    rng = np.random.default_rng(seed)
    # calibrate
    t0 = time.perf_counter(); steps = 0; total = 0.0
    while True:
        t1 = time.perf_counter()
        if with_act:
            a, v, state, nlp, mean = LO.act(P, obs, state, done.astype(np.float64), rng.normal(size=(n_envs, 12)), dtype=np.float32)
            act = np.clip(a, -1, 1).astype(np.float32)
        else:
            act = np.clip(rng.normal(0, 0.1, size=(n_envs, 12)), -1, 1).astype(np.float32)
        obs, rew, done, extra = o.step(act)
        total += time.perf_counter() - t1; steps += 1
        if time.perf_counter() - t0 > budget_s and steps >= 3:
            break
    return n_envs * steps / total, cores, f"{steps} control steps x {n_envs} envs ({'numpy LSTM act + ' if with_act else ''}oracle OpenMP step, double precision)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.cpu_envs
    # every "step" of this arm is one control step over the bounded sample (n envs); W + K steps on the host cores
    from oracle_lib import Oracle
    from oracle import lstm_oracle as LO
    from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES
    cores = os.cpu_count() or 1
    cfg = workload_cfg(n); cfg["num_threads"] = cores
    o = Oracle(cfg)
    W, wname = policy_weights(); P = dict(zip(PARAM_NAMES, W))
    obs = o.reset(); state = np.zeros((n, 384)); done = np.zeros(n, bool)
    rng = np.random.default_rng(0)
    K = min(args.steps, 400); Wm = min(args.warmup, 20)
    total = 0.0
    for it in range(Wm + K):
        t1 = time.perf_counter()
        a, v, state, nlp, mean = LO.act(P, obs, state, done.astype(np.float64), rng.normal(size=(n, 12)), dtype=np.float32)
        obs, rew, done, extra = o.step(np.clip(a, -1, 1).astype(np.float32))
        if it >= Wm:
            total += time.perf_counter() - t1
    val = n * K / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"bp5 trot imitation (LSTM act + env step), bounded sample of {n} envs on the host cores",
                       "envs": n, "policy": wname},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{K} control steps x {n} envs; oracle restatement of the reference's OpenMP step (RaiSim unavailable) + numpy LSTM act"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
    from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import dump_yaml
    from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local); torch.cuda.set_device(dev)
    L = _lib.load()
    N = args.envs_per_gpu; K = args.steps; Wm = max(args.warmup, 3)
    cfg = workload_cfg(N)
    env = FlexibleGymEnv("", dump_yaml(cfg), device=local, env_offset=rank * N)
    # one explicit (non-default) stream carries everything: env/policy kernels, the L2 flush and the timing events
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    env.setStream(stream.cuda_stream)
    env.init()
    W, wname = policy_weights()
    pol = FusedLstmPolicy(W, n_env=N, device=local, seed=1, env_offset=rank * N)

    # ---- device-resident rollout buffers ([K,N,...], written once per step like Runner.run's mb_* lists)
    f32 = dict(device=dev, dtype=torch.float32)
    buf_obs = torch.empty((K, N, 35), **f32); buf_act = torch.empty((K, N, 12), **f32); buf_val = torch.empty((K, N), **f32)
    buf_nlp = torch.empty((K, N), **f32); buf_rew = torch.empty((K, N), **f32); buf_done = torch.empty((K, N), device=dev, dtype=torch.uint8)
    cur_obs = torch.zeros((N, 35), **f32); cur_done = torch.zeros((N,), device=dev, dtype=torch.uint8); state = torch.zeros((N, 384), **f32)
    env.reset(cur_obs)
    flush = None if args.no_l2_flush else torch.empty(256 << 20, device=dev, dtype=torch.uint8)

    def rollout_one(t):
        b = _lib.RolloutBuffers(obs=buf_obs[t].data_ptr(), actions=buf_act[t].data_ptr(), values=buf_val[t].data_ptr(), neglogps=buf_nlp[t].data_ptr(),
                                rewards=buf_rew[t].data_ptr(), dones=buf_done[t].data_ptr(), cur_obs=cur_obs.data_ptr(), cur_done=cur_done.data_ptr(),
                                state=state.data_ptr(), ep_return=None, ep_length=None)
        _lib.check(L.irrl_rollout(env.handle, pol.handle, 1, C.byref(b), 0), "rollout")

    for t in range(Wm):
        rollout_one(t % K)
    torch.cuda.synchronize(dev)
    # ---- timed region: exactly K steps, CUDA events on the launching stream, L2 flushed between steps
    L.irrl_set_profiling(env.handle, 1)
    sampler = ClockSampler(local); sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for t in range(K):
        if flush is not None:
            flush.fill_(t & 0xFF)
        ev[t][0].record(stream)
        rollout_one(t)
        ev[t][1].record(stream)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    act_ms, step_ms, cnt = C.c_double(), C.c_double(), C.c_int64()
    L.irrl_get_profile(env.handle, C.byref(act_ms), C.byref(step_ms), C.byref(cnt))
    L.irrl_set_profiling(env.handle, 0)
    t_ms = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_ms.item())
    value = world * N * K / (total_ms_max * 1e-3)
    sw = np.zeros(N, np.int32); env.getSolverSweeps(sw)
    episodes_done = int(buf_done.sum().item())
    mean_rew = float(buf_rew.mean().item())

    # ---- e2e: the reference-facing calls with HOST buffers (model.step + env.step), copies inside the timed region
    Ke = max(10, min(args.e2e_steps, K))
    # RaisimGymVecEnv / the policy object own their host buffers as one page-locked block each, laid out like the native output
    # blocks (vec_env.py, _lib.pinned_block), so every call returns its results in a single DMA copy
    h_obs, h_rew, h_extra, h_done = _lib.pinned_block(((N, 35), np.float32), ((N,), np.float32), ((N, 6), np.float32), ((N,), np.bool_))
    h_act, h_clip, h_val, h_nlp = _lib.pinned_block(((N, 12), np.float32), ((N, 12), np.float32), ((N,), np.float32), ((N,), np.float32))
    env.reset(h_obs)
    state.zero_()

    act_args = (C.c_void_p(stream.cuda_stream), N, C.c_void_p(h_obs.ctypes.data), C.c_void_p(h_done.view(np.uint8).ctypes.data), C.c_void_p(state.data_ptr()),
                C.c_void_p(h_act.ctypes.data), C.c_void_p(h_clip.ctypes.data), C.c_void_p(h_val.ctypes.data), C.c_void_p(h_nlp.ctypes.data))

    def e2e_step(t):
        # model.step(obs, states, dones): obs/mask host -> device, actions/values/neglogp device -> host; LSTM state stays
        # an opaque device-resident handle (the Runner only threads `states` through, ppo2.py:520)
        _lib.check(L.irrl_policy_act(pol.handle, *act_args, 0, 1, rank * N, 100000 + t), "policy_act")
        env.step(h_clip, h_obs, h_rew, h_done, h_extra)     # RaisimGymVecEnv.step -> wrapper.step (RaisimGymVecEnv.py:31)

    for t in range(3):
        e2e_step(t)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        e2e_step(3 + t)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * N * Ke / float(t_e.item())
    h2d = N * (35 * 4 + 1 + 12 * 4)                        # obs + mask (act)  + clipped action (step)
    d2h = N * (12 * 4 * 2 + 4 + 4 + 35 * 4 + 4 + 1 + 6 * 4)  # action, clipped, value, neglogp | obs, reward, done, extra

    if rank != 0:
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
    step_kernel_ms = step_ms.value / max(cnt.value, 1)
    act_kernel_ms = act_ms.value / max(cnt.value, 1)
    flop_per_env_step = counts.get("env_step", {}).get("fp32_flop_per_env_step")
    traffic = counts.get("env_step", {}).get("dram_bytes_per_launch")
    roof = {"kernel": "env_step_kernel", "bound": "fp32", "unit": "TFLOP/s", "peak": fp32_meas.value or NOMINAL_FP32_TFLOPS,
            "peak_source": "measured live: register-resident FMA micro-benchmark (irrl_measure_fp32_peak); nominal 74.5",
            "kernel_ms": step_kernel_ms, "share_of_step": step_kernel_ms / (total_ms / K), "traffic": traffic,
            "hbm": {"achieved": N * B_STEP_BYTES / (step_kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src,
                    "frac": N * B_STEP_BYTES / (step_kernel_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_env_step": B_STEP_BYTES}}
    if flop_per_env_step:
        ach = N * flop_per_env_step / (step_kernel_ms * 1e-3) / 1e12
        roof["oracle_flop_per_env_step"] = counts.get("oracle", {}).get("flop_per_env_step")   # dense CPU formulation, reported only
        roof.update({"achieved": ach, "frac": ach / roof["peak"], "flop_per_env_step": flop_per_env_step,
                     "flop_source": counts.get("env_step", {}).get("source", "ncu")})
    else:
        roof.update({"achieved": None, "frac": None, "flop_per_env_step": None, "flop_source": "profiles/kernel_counts.json missing"})
    roof["lstm_act"] = {"kernel_ms": act_kernel_ms, "bound": "hbm", "achieved": N * B_ACT_BYTES / (act_kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": N * B_ACT_BYTES / (act_kernel_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_env": B_ACT_BYTES,
                        "kernel": "lstm_act_tc_kernel: tcgen05.mma kind::tf32 x3 split, TMEM accumulators, cp.async.bulk operand/state staging" if N >= 256 else "lstm_act_kernel (fp32 FMA)",
                        "tensor_tflops_issued": N * 2 * 3 * 2 * (88 * 192 + 96 * 192) / (act_kernel_ms * 1e-3) / 1e12 if N >= 256 else None}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": total_ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": {"trot": f"bp5 trot imitation reward, {N} envs per B200 (BASELINE.json configs[1])",
                                     "relaxation": f"bp5 relaxation phase (mimic reward removed), {N} envs per B200 (BASELINE.json configs[2])",
                                     "stairs": f"bp5 stair heightfield + domain randomisation, {N} envs per B200 (BASELINE.json configs[4])"}[WORKLOAD]
                                    + " with fused LSTM act: act kernel + env step kernel (PD, 8 physics substeps, obs, reward, done, auto-reset) + rollout stores",
                       "envs_per_gpu": N, "total_envs": world * N, "substeps": 8, "policy": wname, "obs_noise": 2.0, "domain_randomisation": True,
                       "l2": "flushed between timed steps (256 MiB write)" if flush is not None else "not flushed (state << L2)",
                       "parallelism": f"env-shard x{world}, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "path": "irrl_policy_act + irrl_step with page-locked host numpy buffers (as RaisimGymVecEnv owns them: one block per call, one DMA copy back), 2 blocking calls per step; LSTM state device-resident"},
            "gpu_launches": 2 * K, "kernels": ["lstm_act_tc_kernel (tcgen05 3xTF32)" if N >= 256 else "lstm_act_kernel", "env_step_kernel"], "memcpy_d2d_per_step": 0,
            "roofline": roof, "clocks": clocks, "wall_s": wall,
            "sanity": {"mean_reward": mean_rew, "episodes_finished": episodes_done, "gs_sweeps_last_substep": {"mean": float(sw.mean()), "max": int(sw.max()), "hist": np.bincount(sw, minlength=9).tolist()}}}
    if not args.no_cpu_baseline:
        v, cores, sample = cpu_arm(args.cpu_envs, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    global WORKLOAD
    args = parse()
    WORKLOAD = args.workload
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
