"""CPU: the C-ABI library loads, exports every symbol include/irrl_b200.h declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "irrl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(irrl_[a-z_0-9]+)\s*\(", text)))


def _lib():
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported():
    L = _lib()
    names = _declared()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "irrl_b200.h"\nint main(void){ irrl_rollout_buffers b; (void)b; return IRRL_OB_DIM == 35 ? 0 : 1; }\n')
    import subprocess
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib()
    h = C.c_void_p()
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, dump_yaml
    rc = L.irrl_create(b"", dump_yaml(trot_cfg()).encode(), 0, 0, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in L.irrl_last_error()
    from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
    with pytest.raises(RuntimeError):
        FlexibleGymEnv("", dump_yaml(trot_cfg()))


@pytest.mark.gpu
def test_yaml_missing_key_is_reported_like_read_yaml():
    """GYM:41-42 READ_YAML aborts with "Node ... doesn't exist"; here the constructor raises with the same message."""
    from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv
    from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg, dump_yaml
    cfg = trot_cfg(); del cfg["Stiffness"]
    with pytest.raises(RuntimeError, match="Stiffness"):
        FlexibleGymEnv("", dump_yaml(cfg))


def test_pinned_block_lays_arrays_out_back_to_back():
    """_lib.pinned_block carves one block into the arrays of a C-ABI call in the order the native side writes them (one DMA copy per
    call); without a GPU the block simply stays pageable"""
    import numpy as np
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
    n = 37
    ob, rew, extra, done = _lib.pinned_block(((n, 35), np.float32), ((n,), np.float32), ((n, 6), np.float32), ((n,), np.bool_))
    assert ob.shape == (n, 35) and rew.shape == (n,) and extra.shape == (n, 6) and done.dtype == np.bool_
    assert rew.ctypes.data == ob.ctypes.data + ob.nbytes and extra.ctypes.data == rew.ctypes.data + rew.nbytes and done.ctypes.data == extra.ctypes.data + extra.nbytes
    ob[:] = 1; rew[:] = 2; extra[:] = 3; done[:] = True
    assert ob.sum() == n * 35 and rew.sum() == 2 * n and extra.sum() == 18 * n and done.all()
