"""`_flexible_robot.FlexibleGymEnv` -- the class the reference binds with pybind11 (raisim_gym.cpp:14-47) over
VectorizedEnvironment<ENVIRONMENT> (VectorizedEnvironment.hpp:127-382), here a thin ctypes shim over the C ABI
(include/irrl_b200.h) of the CUDA library.  Same constructor, same 29 methods, same in-place numpy semantics:
arrays must be C-contiguous float32 (bool for `done`) of the right shape, otherwise TypeError -- pybind's
Eigen::Ref refuses to copy writable arguments.

Additions (not in the reference): torch CUDA tensors are accepted wherever a numpy array is (zero-copy, asynchronous
on the env's stream), `device=`/`env_offset=` keyword arguments for multi-GPU sharding, and state get/set hooks.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from . import _lib

__all__ = ["FlexibleGymEnv"]


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class FlexibleGymEnv:
    def __init__(self, resourceDir: str, cfg: str, device: int = 0, env_offset: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._L.irrl_create(str(resourceDir).encode(), str(cfg).encode(), int(device), int(env_offset), C.byref(self._h)),
                   "FlexibleGymEnv")
        self._n = self._L.irrl_get_num_envs(self._h)
        self.device = int(device)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.irrl_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- argument checking in the spirit of pybind11's Eigen::Ref<..., RowMajor> (no implicit copies)
    def _ptr(self, a, shape, dtype=np.float32, name="array"):
        if _is_torch(a):
            import torch
            want = {np.float32: torch.float32, np.bool_: torch.bool, np.uint8: torch.uint8, np.int32: torch.int32}[dtype]
            ok_dtype = a.dtype == want or (dtype == np.bool_ and a.dtype == torch.uint8)
            if not (a.is_cuda and a.is_contiguous() and ok_dtype and tuple(a.shape) == tuple(shape)):
                raise TypeError(f"{name}: expected a contiguous CUDA tensor {dtype.__name__}{list(shape)}, got {a.dtype}{list(a.shape)} on {a.device}")
            return C.c_void_p(a.data_ptr())
        if not isinstance(a, np.ndarray):
            raise TypeError(f"{name}: expected numpy.ndarray, got {type(a).__name__}")
        ok_dtype = a.dtype == dtype or (dtype == np.bool_ and a.dtype == np.uint8)
        if not ok_dtype or not a.flags["C_CONTIGUOUS"] or tuple(a.shape) != tuple(shape):
            raise TypeError(f"{name}: incompatible function arguments: expected C-contiguous {np.dtype(dtype).name}{list(shape)}, "
                            f"got {a.dtype}{list(a.shape)} (contiguous={a.flags['C_CONTIGUOUS']})")
        return C.c_void_p(a.ctypes.data)

    # ---- raisim_gym.cpp:17-46
    def init(self) -> None:
        _lib.check(self._L.irrl_init(self._h), "init")

    def getExtraInfoNames(self) -> List[str]:
        return [self._L.irrl_get_extra_info_name(self._h, i).decode() for i in range(self._L.irrl_get_extra_info_dim(self._h))]

    def reset(self, ob) -> None:
        _lib.check(self._L.irrl_reset(self._h, self._ptr(ob, (self._n, 35), name="ob")), "reset")

    def observe(self, ob) -> None:
        _lib.check(self._L.irrl_observe(self._h, self._ptr(ob, (self._n, 35), name="ob")), "observe")

    def step(self, action, ob, reward, done, extraInfo) -> None:
        _lib.check(self._L.irrl_step(self._h, self._ptr(action, (self._n, 12), name="action"), self._ptr(ob, (self._n, 35), name="ob"),
                                     self._ptr(reward, (self._n,), name="reward"), self._ptr(done, (self._n,), np.bool_, name="done"),
                                     self._ptr(extraInfo, (self._n, 6), name="extraInfo")), "step")

    def testStep(self, action, ob, reward, done, extraInfo) -> None:
        _lib.check(self._L.irrl_test_step(self._h, self._ptr(action, (self._n, 12), name="action"), self._ptr(ob, (self._n, 35), name="ob"),
                                          self._ptr(reward, (self._n,), name="reward"), self._ptr(done, (self._n,), np.bool_, name="done"),
                                          self._ptr(extraInfo, (self._n, 6), name="extraInfo")), "testStep")

    def setSeed(self, seed: int) -> None:
        _lib.check(self._L.irrl_set_seed(self._h, int(seed)), "setSeed")

    def close(self) -> None:
        _lib.check(self._L.irrl_close(self._h), "close")

    def isTerminalState(self, terminalState) -> None:
        _lib.check(self._L.irrl_is_terminal_state(self._h, self._ptr(terminalState, (self._n,), np.bool_, name="terminalState")), "isTerminalState")

    def setSimulationTimeStep(self, dt: float) -> None:
        _lib.check(self._L.irrl_set_simulation_time_step(self._h, float(dt)))

    def setControlTimeStep(self, dt: float) -> None:
        _lib.check(self._L.irrl_set_control_time_step(self._h, float(dt)))

    def getObDim(self) -> int:
        return self._L.irrl_get_ob_dim(self._h)

    def getActionDim(self) -> int:
        return self._L.irrl_get_action_dim(self._h)

    def getExtraInfoDim(self) -> int:
        return self._L.irrl_get_extra_info_dim(self._h)

    def getNumOfEnvs(self) -> int:
        return self._n

    def startRecordingVideo(self, fileName: str) -> None:
        _lib.check(self._L.irrl_start_recording_video(self._h, str(fileName).encode()))

    def stopRecordingVideo(self) -> None:
        _lib.check(self._L.irrl_stop_recording_video(self._h))

    def showWindow(self) -> None:
        _lib.check(self._L.irrl_show_window(self._h))

    def hideWindow(self) -> None:
        _lib.check(self._L.irrl_hide_window(self._h))

    def curriculumUpdate(self) -> None:
        _lib.check(self._L.irrl_curriculum_update(self._h))

    def OriginState(self, ob) -> None:
        _lib.check(self._L.irrl_origin_state(self._h, self._ptr(ob, (self._n, 41), name="origin_state")), "OriginState")

    def GetOriginStateDim(self) -> int:
        return self._L.irrl_get_origin_state_dim(self._h)

    def ReferenceState(self, refer) -> None:
        # the reference's vector wrapper calls OriginState here by mistake (VectorizedEnvironment.hpp:223-226);
        # this follows the evident intent (Environment.hpp:1339-1345): jointRef | jointDotRef
        _lib.check(self._L.irrl_reference_state(self._h, self._ptr(refer, (self._n, 24), name="refer_state")), "ReferenceState")

    def GetJointEffort(self, joint_effort) -> None:
        _lib.check(self._L.irrl_get_joint_effort(self._h, self._ptr(joint_effort, (self._n, 12), name="joint_effort")), "GetJointEffort")

    def GetGeneralizedForce(self, generalized_force) -> None:
        _lib.check(self._L.irrl_get_generalized_force(self._h, self._ptr(generalized_force, (self._n, 18), name="generalized_force")), "GetGeneralizedForce")

    def GetInverseMassMatrix(self, inverse_mass) -> None:
        _lib.check(self._L.irrl_get_inverse_mass_matrix(self._h, self._ptr(inverse_mass, (self._n, 324), name="inverse_mass")), "GetInverseMassMatrix")

    def GetNonlinear(self, nonlinear) -> None:
        _lib.check(self._L.irrl_get_nonlinear(self._h, self._ptr(nonlinear, (self._n, 18), name="nonlinear")), "GetNonlinear")

    def SetContactCoefficient(self, contact_coeff) -> None:
        _lib.check(self._L.irrl_set_contact_coefficient(self._h, self._ptr(contact_coeff, (self._n, 3), name="contact_coeff")), "SetContactCoefficient")

    def GetSphereInfo(self, sphere_info) -> None:
        _lib.check(self._L.irrl_get_sphere_info(self._h, self._ptr(sphere_info, (self._n, 4), name="sphere_info")), "GetSphereInfo")

    # ---- additions
    def getMeteor(self, out) -> None:
        """[N,9] float32 host array: sphere position, velocity, mode (0 placed / 1 falling), radius, mass (Crutial: True)"""
        _lib.check(self._L.irrl_get_meteor(self._h, self._ptr(out, (self._n, 9), name="meteor")), "getMeteor")

    def setMeteor(self, m) -> None:
        _lib.check(self._L.irrl_set_meteor(self._h, self._ptr(m, (self._n, 9), name="meteor")), "setMeteor")

    def GetMassMatrix(self, mass) -> None:
        _lib.check(self._L.irrl_get_mass_matrix(self._h, self._ptr(mass, (self._n, 324), name="mass")), "GetMassMatrix")

    def getState(self, out) -> None:
        _lib.check(self._L.irrl_get_state(self._h, self._ptr(out, (self._n, 192), name="state")), "getState")

    def setState(self, state) -> None:
        _lib.check(self._L.irrl_set_state(self._h, self._ptr(state, (self._n, 192), name="state")), "setState")

    def getModelParams(self, out) -> None:
        _lib.check(self._L.irrl_get_model_params(self._h, self._ptr(out, (self._n, 94), name="model_params")), "getModelParams")

    def integrate(self, tau, contact_out) -> None:
        _lib.check(self._L.irrl_integrate(self._h, self._ptr(tau, (self._n, 12), name="tau"), self._ptr(contact_out, (self._n, 16), name="contact_out")), "integrate")

    def getSolverSweeps(self, out) -> None:
        _lib.check(self._L.irrl_get_solver_sweeps(self._h, self._ptr(out, (self._n,), np.int32, name="sweeps")), "getSolverSweeps")

    def setRefTraj(self, table) -> None:
        table = np.ascontiguousarray(table, np.float32)
        if table.ndim != 2 or table.shape[1] != 30:
            raise TypeError("reference table must be [rows,30] (Environment.hpp:17-21)")
        _lib.check(self._L.irrl_set_ref_traj(self._h, C.c_void_p(table.ctypes.data), int(table.shape[0])), "setRefTraj")

    def setHeightfield(self, heights, x_size: float, y_size: float, cx: float = 0.0, cy: float = 0.0) -> None:
        heights = np.ascontiguousarray(heights, np.float32)
        if heights.ndim != 2:
            raise TypeError("heights must be [nx, ny]")
        _lib.check(self._L.irrl_set_heightfield(self._h, C.c_void_p(heights.ctypes.data), int(heights.shape[0]), int(heights.shape[1]),
                                                float(x_size), float(y_size), float(cx), float(cy)), "setHeightfield")

    def generateTerrain(self, kind: str = "perlin", nx: int = 5000, ny: int = 500, x_size: float = 500.0, y_size: float = 20.0) -> None:
        _lib.check(self._L.irrl_generate_terrain(self._h, kind.encode(), int(nx), int(ny), float(x_size), float(y_size)), "generateTerrain")

    def getHeightfield(self):
        nx, ny = C.c_int(), C.c_int(); xs, ys, cx, cy = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        _lib.check(self._L.irrl_get_heightfield(self._h, None, C.byref(nx), C.byref(ny), C.byref(xs), C.byref(ys), C.byref(cx), C.byref(cy)), "getHeightfield")
        h = np.zeros((nx.value, ny.value), np.float32)
        _lib.check(self._L.irrl_get_heightfield(self._h, C.c_void_p(h.ctypes.data), None, None, None, None, None, None), "getHeightfield")
        return h, xs.value, ys.value, cx.value, cy.value

    def lastEpisodeStats(self, ep_return, ep_length) -> None:
        _lib.check(self._L.irrl_last_episode_stats(self._h, self._ptr(ep_return, (self._n,), name="ep_return"),
                                                   self._ptr(ep_length, (self._n,), np.int32, name="ep_length")), "lastEpisodeStats")

    def runningEpisodeStats(self, ep_return, ep_length, clear: bool = False) -> None:
        _lib.check(self._L.irrl_running_episode_stats(self._h, self._ptr(ep_return, (self._n,), name="ep_return"),
                                                      self._ptr(ep_length, (self._n,), np.int32, name="ep_length"), int(clear)), "runningEpisodeStats")

    def setTick(self, tick: int) -> None:
        _lib.check(self._L.irrl_set_tick(self._h, int(tick)))

    def getTick(self) -> int:
        return int(self._L.irrl_get_tick(self._h))

    def setStream(self, cuda_stream_ptr: int) -> None:
        _lib.check(self._L.irrl_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    @property
    def handle(self):
        return self._h
