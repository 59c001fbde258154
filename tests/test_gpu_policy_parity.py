"""-m gpu: fused LSTM act kernels (fp32 FMA kernel below 256 environments, tcgen05 3xTF32 kernel from 256) and GAE
kernel against the numpy oracle (oracle/lstm_oracle.py, itself pinned on the reference's CustomerLstmNN outputs) with
the trained bp5_155 weights.  fp32 tolerance: 2e-5 absolute on action means / states (O(1) numbers), 2e-5 relative to
the largest |value| on the value head (values reach ~30)."""
import ctypes as C
import os

import numpy as np
import pytest

from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy, PARAM_NAMES, init_params
from oracle import lstm_oracle as LO
from oracle_lib import lib as oracle_lib

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5


def _weights():
    z = np.load(os.path.join(G, "bp5_155_params.npz"))
    return [z[k] for k in PARAM_NAMES]


def _eps(seed, env_ids, tick):
    """the kernel's Gaussian draws: Philox block (seed, env, tick, 128 + a//4) -> Box-Muller, word a%4"""
    L = oracle_lib()
    out = np.zeros((len(env_ids), 12), np.float32)
    r = (C.c_uint * 4)()
    for i, e in enumerate(env_ids):
        for blk in range(3):
            L.bp5o_philox(C.c_uint(seed), C.c_uint(int(e)), C.c_uint(tick), C.c_uint(128 + blk), r)
            w = [int(x) for x in r]
            for p in range(2):
                u1 = np.float32(((w[2 * p] >> 8) + 1)) * np.float32(1.0 / 16777216.0)
                u2 = np.float32(w[2 * p + 1] >> 8) * np.float32(1.0 / 16777216.0)
                rad = np.sqrt(-2.0 * np.log(np.float64(u1))); ang = 6.283185307179586 * np.float64(u2)
                out[i, 4 * blk + 2 * p] = rad * np.cos(ang); out[i, 4 * blk + 2 * p + 1] = rad * np.sin(ang)
    return out


@pytest.mark.parametrize("n", [1, 16, 100, 1000])
def test_act_matches_numpy_oracle(n):
    W = _weights(); P = dict(zip(PARAM_NAMES, W))
    pol = FusedLstmPolicy(W, n_env=n, seed=7, env_offset=3)
    rng = np.random.default_rng(n)
    obs = rng.normal(0, 0.7, size=(n, 35)).astype(np.float32)
    state = (rng.normal(0, 0.4, size=(n, 384))).astype(np.float32)
    mask = (rng.random(n) < 0.3)
    # deterministic
    a, v, s, nlp = pol.step(obs, state, mask, deterministic=True, tick=11)
    ra, rv, rs, rnlp, rmean = LO.act(P, obs, state, mask.astype(np.float64))
    assert np.abs(a - rmean).max() < TOL and np.abs(v - rv).max() < TOL * max(1.0, np.abs(rv).max()) and np.abs(s - rs).max() < TOL
    assert np.abs(nlp - rnlp).max() < 1e-4
    # stochastic with the same counter-based draws
    a, v, s, nlp, clip = pol.step(obs, state, mask, deterministic=False, tick=12, return_clipped=True)
    eps = _eps(7, np.arange(n) + 3, 12)
    ra, rv, rs, rnlp, rmean = LO.act(P, obs, state, mask.astype(np.float64), eps)
    assert np.abs(a - ra).max() < 5e-5
    assert np.abs(nlp - rnlp).max() < 2e-3 * max(1.0, np.abs(rnlp).max())
    assert np.abs(clip - np.clip(a, -1, 1)).max() == 0
    assert np.abs(s - rs).max() < TOL


@pytest.mark.parametrize("n", [128, 300, 301, 4096])     # 301: last tile of 45 rows -> observation tile not 16-byte granular (plain-load path)
def test_tensor_core_kernel_matches_fma_kernel_and_oracle(n):
    """both act kernels forced on the same inputs (irrl_policy_set_act_path): same Philox draws, same outputs"""
    L = _lib.load()
    W = _weights(); P = dict(zip(PARAM_NAMES, W))
    rng = np.random.default_rng(n)
    obs = rng.normal(0, 0.7, size=(n, 35)).astype(np.float32); state = rng.normal(0, 0.4, size=(n, 384)).astype(np.float32)
    mask = rng.random(n) < 0.3
    pol = FusedLstmPolicy(W, n_env=n, seed=7, env_offset=3)
    out = {}
    try:
        for mode in (1, 2):
            _lib.check(L.irrl_policy_set_act_path(mode))
            out[mode] = pol.step(obs, state, mask, deterministic=False, tick=12, return_clipped=True)
            out[mode + 10] = pol.step(obs, state, mask, deterministic=True, tick=12)
    finally:
        _lib.check(L.irrl_policy_set_act_path(0))
    ra, rv, rs, rnlp, rmean = LO.act(P, obs, state, mask.astype(np.float64))
    vs = max(1.0, np.abs(rv).max())
    for i, tol in enumerate((1e-5, 1e-5 * vs, 1e-5, 1e-4, 1e-5)):      # action, value, state, neglogp, clipped
        assert np.abs(out[1][i] - out[2][i]).max() < tol, (i, np.abs(out[1][i] - out[2][i]).max())
    for mode in (11, 12):
        a, v, s, nlp = out[mode]
        assert np.abs(a - rmean).max() < TOL and np.abs(v - rv).max() < TOL * vs and np.abs(s - rs).max() < TOL and np.abs(nlp - rnlp).max() < 1e-4


def test_tcgen05_gemm_probe_3xtf32_accuracy():
    """descriptor / TMEM path on its own: 3xTF32 product is fp32-accurate, a single tf32 pass is ~1e-3 (proves the split is live)"""
    L = _lib.load()
    rng = np.random.default_rng(5)
    for k, n in [(8, 16), (32, 192), (48, 16), (64, 192)]:
        a = rng.normal(size=(128, k)).astype(np.float32); b = rng.normal(size=(n, k)).astype(np.float32)
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        d3 = np.zeros((128, n), np.float32); d1 = np.zeros((128, n), np.float32)
        _lib.check(L.irrl_tc_gemm_probe(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_void_p(d3.ctypes.data), k, n, 0))
        _lib.check(L.irrl_tc_gemm_probe(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_void_p(d1.ctypes.data), k, n, 1))
        scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64).T
        assert (np.abs(d3 - ref) / scale).max() < 2e-6
        assert 1e-5 < (np.abs(d1 - ref) / scale).max() < 2e-3


def test_reference_known_answer_sequence_on_gpu():
    """the 64-step sequence the reference's own numpy policy produced (tests/golden/lstm_kat.npz)"""
    W = _weights()
    k = np.load(os.path.join(G, "lstm_kat.npz"))
    pol = FusedLstmPolicy(W, n_env=1)
    state = np.zeros((1, 384), np.float32)
    for t, x in enumerate(k["seq"]):
        a, v, state, nlp = pol.step(x[None].astype(np.float32), state, np.zeros(1, bool), deterministic=True)
        assert np.abs(np.clip(a[0], -1, 1) - k["pkl_action"][t]).max() < 5e-5, t
        assert abs(v[0] - k["pkl_value"][t]) < 5e-5
    assert np.abs(state[0, :192] - k["pkl_state_pi"][-1]).max() < 5e-5
    assert np.abs(state[0, 192:] - k["pkl_state_v"][-1]).max() < 5e-5
    expect = np.array([0.04697, -0.24707, 0.04621, 0.01774, 0.18719, 0.10385, 0.11012, 0.12886, 0.07297, -0.07416, -0.30378, -0.06146])
    a0 = pol.step(k["seq"][:1].astype(np.float32), np.zeros((1, 384), np.float32), None, deterministic=True)[0]
    assert np.abs(np.clip(a0[0], -1, 1) - expect).max() < 3e-5


def test_random_init_weights_and_state_roundtrip():
    W = init_params(np.random.default_rng(3)); P = dict(zip(PARAM_NAMES, W))
    n = 257
    pol = FusedLstmPolicy(W, n_env=n)
    rng = np.random.default_rng(5)
    state = np.zeros((n, 384), np.float32); rstate = np.zeros((n, 384))
    for t in range(5):
        obs = rng.normal(size=(n, 35)).astype(np.float32)
        mask = rng.random(n) < 0.1
        a, v, state, nlp = pol.step(obs, state, mask, deterministic=True)
        ra, rv, rstate, rnlp, rmean = LO.act(P, obs, rstate, mask.astype(np.float64))
        assert np.abs(a - rmean).max() < 5e-5 and np.abs(state - rstate).max() < 5e-5 and np.abs(v - rv).max() < 5e-5


def test_gae_kernel_matches_loop():
    import torch
    L = _lib.load()
    rng = np.random.default_rng(9)
    T, N = 750, 300
    r = rng.normal(size=(T, N)).astype(np.float32); v = rng.normal(size=(T, N)).astype(np.float32)
    d = (rng.random((T, N)) < 0.01); lv = rng.normal(size=N).astype(np.float32); ld = (rng.random(N) < 0.01)
    adv, ret = LO.gae(r.astype(np.float64), v.astype(np.float64), d.astype(np.float64), lv.astype(np.float64), ld.astype(np.float64), 0.99, 0.998)
    dev = torch.device("cuda:0")
    tr, tv = torch.from_numpy(r).to(dev), torch.from_numpy(v).to(dev)
    td, tlv, tld = torch.from_numpy(d.astype(np.uint8)).to(dev), torch.from_numpy(lv).to(dev), torch.from_numpy(ld.astype(np.uint8)).to(dev)
    ta, tret = torch.empty_like(tr), torch.empty_like(tr)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.irrl_gae(C.c_void_p(st), T, N, C.c_void_p(tr.data_ptr()), C.c_void_p(tv.data_ptr()), C.c_void_p(td.data_ptr()), C.c_void_p(tlv.data_ptr()),
                          C.c_void_p(tld.data_ptr()), C.c_float(0.99), C.c_float(0.998), C.c_void_p(ta.data_ptr()), C.c_void_p(tret.data_ptr())))
    torch.cuda.synchronize()
    assert np.abs(ta.cpu().numpy() - adv).max() < 2e-3 * np.abs(adv).max()
    assert np.abs(tret.cpu().numpy() - ret).max() < 2e-3 * np.abs(ret).max()
