"""In-tree build of the CUDA library (nvcc, sm_100a only).  `python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libirrl_b200.so")
SOURCES = ["env_kernels.cu", "policy_kernels.cu", "policy_tc_kernels.cu", "lstm_seq_mma.cu", "learner_gemm.cu", "learner_tc.cu", "ppo_head_loss.cu", "capi.cu"]
HEADERS = ["env_device.cuh", "env_kernels.h", "irrl_params.h", "tc_common.cuh", os.path.join("..", "..", "include", "irrl_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
if os.environ.get("IRRL_FASTDIV"):
    NVCC_FLAGS += ["-DIRRL_FASTDIV=" + os.environ["IRRL_FASTDIV"]]
if os.environ.get("IRRL_SINCOS"):               # 0 libdevice sincosf, 1 MUFU, 2 (default) Cody-Waite polynomial
    NVCC_FLAGS += ["-DIRRL_SINCOS=" + os.environ["IRRL_SINCOS"]]
if os.environ.get("IRRL_EXP"):                  # experiment switches: space-separated -D macros
    NVCC_FLAGS += ["-D" + m for m in os.environ["IRRL_EXP"].split()]
if os.environ.get("IRRL_STEP_MINWARPS"):       # tuning knob: register cap of the step kernel = 65536 / (32 * resident warps per SM)
    NVCC_FLAGS += ["-DSTEP_MINWARPS=" + os.environ["IRRL_STEP_MINWARPS"]]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    deps = [os.path.join(CSRC, h) for h in HEADERS]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + deps):
            jobs.append([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            print(l)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"])
    build_pybind(force=force or bool(jobs))
    return LIB


PYMOD = os.path.join(HERE, "_flexible_robot_pb.so")


def build_pybind(force: bool = False) -> str:
    """the compiled pybind11 module `_flexible_robot_pb` (csrc/pybind_shim.cpp): class FlexibleGymEnv as the reference binds it
    (raisim_gym.cpp:14-47), linked against libirrl_b200.so next to it ($ORIGIN rpath)."""
    import sysconfig
    src = os.path.join(CSRC, "pybind_shim.cpp")
    hdr = os.path.join(CSRC, "..", "..", "include", "irrl_b200.h")
    if not force and not _stale(PYMOD, [src, hdr, LIB]):
        return PYMOD
    import pybind11
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
           src, "-o", PYMOD, "-L" + HERE, "-l:libirrl_b200.so", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("pybind module build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return PYMOD


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
