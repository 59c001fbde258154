"""ctypes binding of the C ABI (include/irrl_b200.h).  There is no CPU fallback: if the shared library is missing
or no CUDA device is present every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libirrl_b200.so")
_lib = None

c_f32p = C.c_void_p   # raw addresses: host numpy buffers or device pointers


class RolloutBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "actions", "values", "neglogps", "rewards", "dones", "cur_obs", "cur_done",
                                           "state", "ep_return", "ep_length")]


class ActStepIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "done", "state", "action", "clipped", "value", "neglogp", "next_obs", "reward", "next_done", "extra")]


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m high_speed_quadrupedal_locomotion_by_irrl_b200.build` "
                           "(the CUDA extension is mandatory, there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.irrl_last_error.restype = C.c_char_p
    L.irrl_version.restype = C.c_char_p
    L.irrl_get_extra_info_name.restype = C.c_char_p
    L.irrl_get_extra_info_name.argtypes = [C.c_void_p, C.c_int]
    L.irrl_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.irrl_destroy.argtypes = [C.c_void_p]
    L.irrl_destroy.restype = None
    L.irrl_get_tick.restype = C.c_uint32
    L.irrl_get_tick.argtypes = [C.c_void_p]
    L.irrl_set_tick.argtypes = [C.c_void_p, C.c_uint32]
    for name in ("init", "close", "curriculum_update", "stop_recording_video", "show_window", "hide_window",
                 "get_ob_dim", "get_action_dim", "get_extra_info_dim", "get_num_envs", "get_origin_state_dim"):
        getattr(L, "irrl_" + name).argtypes = [C.c_void_p]
    for name in ("reset", "observe", "is_terminal_state", "origin_state", "reference_state", "get_joint_effort", "get_generalized_force",
                 "get_inverse_mass_matrix", "get_nonlinear", "get_mass_matrix", "set_contact_coefficient", "get_sphere_info",
                 "get_model_params", "get_state", "set_state", "get_solver_sweeps", "set_stream", "get_meteor", "set_meteor"):
        getattr(L, "irrl_" + name).argtypes = [C.c_void_p, C.c_void_p]
    L.irrl_step.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.irrl_test_step.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.irrl_last_episode_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.irrl_running_episode_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.irrl_set_seed.argtypes = [C.c_void_p, C.c_int]
    L.irrl_set_simulation_time_step.argtypes = [C.c_void_p, C.c_double]
    L.irrl_set_control_time_step.argtypes = [C.c_void_p, C.c_double]
    L.irrl_start_recording_video.argtypes = [C.c_void_p, C.c_char_p]
    L.irrl_integrate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.irrl_set_ref_traj.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.irrl_policy_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.irrl_policy_set_params.argtypes = [C.c_void_p, C.c_void_p]
    L.irrl_policy_destroy.argtypes = [C.c_void_p]
    L.irrl_policy_destroy.restype = None
    L.irrl_policy_act.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    L.irrl_policy_set_act_path.argtypes = [C.c_int]
    L.irrl_parse_urdf.argtypes = [C.c_char_p, C.c_void_p]
    L.irrl_tc_timeline.argtypes = [C.c_int, C.c_void_p]
    L.irrl_tc_mma_rate.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
    L.irrl_tc_gemm_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.irrl_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(RolloutBuffers), C.c_int]
    L.irrl_act_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(ActStepIO), C.c_int, C.c_uint32, C.c_int]
    L.irrl_gae.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.irrl_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.irrl_get_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.irrl_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.irrl_lstm_seq_fwd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 10
    L.irrl_lstm_seq_bwd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8
    L.irrl_lstm_seq_ctas.argtypes = [C.c_int]
    L.irrl_lstm_seq_set_path.argtypes = [C.c_int]
    L.irrl_proj_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.irrl_proj_rows_set_path.argtypes = [C.c_int]
    L.irrl_ppo_head_loss_ctas.argtypes = [C.c_int, C.c_int]
    L.irrl_scale_unless_one.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    L.irrl_ppo_head_loss.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 11 + [C.c_float, C.c_float, C.c_float] + [C.c_void_p] * 3
    L.irrl_gram2_rows_ctas.argtypes = [C.c_int, C.c_int, C.c_int]
    L.irrl_gram2_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.irrl_gram_rows_ctas.argtypes = [C.c_int, C.c_int, C.c_int]
    L.irrl_gram_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.irrl_lstm_pw_fwd.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8
    L.irrl_lstm_pw_bwd.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 9
    L.irrl_set_heightfield.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    L.irrl_generate_terrain.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double]
    L.irrl_get_heightfield.argtypes = [C.c_void_p, C.c_void_p] + [C.c_void_p] * 6
    L.irrl_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.irrl_host_unregister.argtypes = [C.c_void_p]
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().irrl_last_error().decode(errors="replace")
        raise RuntimeError(f"{what or 'irrl'} failed ({rc}): {msg}")


def pin(array) -> bool:
    """Page-lock a numpy array in place (cudaHostRegister) so the C ABI can DMA straight into it; the registration is
    dropped again when the array is garbage-collected.  Returns False (array stays pageable) when not possible."""
    import weakref
    try:
        L = load()
        addr = array.ctypes.data
        if L.irrl_host_register(C.c_void_p(addr), C.c_size_t(array.nbytes)) != 0:
            return False
        weakref.finalize(array, L.irrl_host_unregister, C.c_void_p(addr))
        return True
    except Exception:
        return False


def pinned_block(*specs):
    """One page-locked block carved into back-to-back arrays: `specs` = (shape, dtype) pairs, each array 4-byte aligned by
    construction when the leading ones are float32.  The C ABI recognises buffers laid out like its device-side output
    blocks ([ob | reward | extra | done] for irrl_step, [action | clipped | value | neglogp] for irrl_policy_act) and moves
    them with a single DMA copy.  Falls back to separate pageable arrays when page-locking is impossible."""
    import numpy as np
    sizes = [int(np.prod(shape)) * np.dtype(dt).itemsize for shape, dt in specs]
    block = np.zeros(sum(sizes), np.uint8)
    pin(block)
    out, off = [], 0
    for (shape, dt), nb in zip(specs, sizes):
        out.append(block[off:off + nb].view(dt).reshape(shape))
        off += nb
    return out
