"""The PPO learner's tensor-core kernels (ppo2.py:136-197 re-hosted): sequence-persistent BPTT with the recurrent products as
mma.sync tf32 3xTF32 MMAs (csrc/lstm_seq_mma.cu) and the streaming products irrl_proj_rows / irrl_gram_rows (csrc/learner_gemm.cu),
all through the C ABI, against float64 autograd of the same recurrence and against the FP32-FMA kernels (the regression reference).
Tolerances: 3xTF32 is fp32-grade -- a few 1e-7 of the tensor's scale per product, 3e-6 is the bar written here."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup():
    import torch
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
    return torch, _lib, _lib.load(), torch.device("cuda:0")


def _seq_inputs(torch, dev, T, K, N, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    r = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
    return dict(xw=r(T, K, N, 192), wh=r(K, 48, 192, sc=0.15), b=r(K, 192, sc=0.1), c0=r(K, N, 48, sc=0.5), h0=r(K, N, 48, sc=0.3),
                keep=(torch.rand(T, N, device=dev, generator=g) > 0.08).float(), dH=r(T, K, N, 48, sc=0.1))


def _run_seq(torch, _lib, L, dev, path, d, T, K, N):
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    prev = L.irrl_lstm_seq_set_path(path)
    try:
        gates = torch.empty(T, K, N, 192, device=dev); Cs = torch.empty(T, K, N, 48, device=dev); Hs = torch.empty_like(Cs); HM = torch.empty_like(Cs)
        _lib.check(L.irrl_lstm_seq_fwd(st, T, K, N, p(d["xw"]), p(d["wh"]), p(d["c0"]), p(d["h0"]), p(d["keep"]), p(gates), p(Cs), p(Hs), p(d["b"]), p(HM)))
        dz = torch.empty_like(gates); db = torch.empty(L.irrl_lstm_seq_ctas(N), K, 192, device=dev)
        _lib.check(L.irrl_lstm_seq_bwd(st, T, K, N, p(d["dH"]), p(d["wh"]), p(d["c0"]), p(d["keep"]), p(gates), p(Cs), p(dz), p(db)))
        torch.cuda.synchronize()
    finally:
        L.irrl_lstm_seq_set_path(prev)
    return dict(gates=gates, Cs=Cs, Hs=Hs, HM=HM, dz=dz, db=db.sum(0))


def _seq_reference(torch, d, T, K, N):
    xw = d["xw"].double().requires_grad_(True); wh = d["wh"].double(); b = d["b"].double().requires_grad_(True)
    c, h = d["c0"].double(), d["h0"].double(); out, cs, hm = [], [], []
    for t in range(T):
        k = d["keep"][t].double().view(1, N, 1); c = c * k; h = h * k; hm.append(h)
        z = xw[t] + b.view(K, 1, 192) + torch.bmm(h, wh)
        i, f, o, g = z.chunk(4, dim=2)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g); h = torch.sigmoid(o) * torch.tanh(c); out.append(h); cs.append(c)
    Hs = torch.stack(out, 0)
    (Hs * d["dH"].double()).sum().backward()
    return dict(Hs=Hs.detach(), Cs=torch.stack(cs, 0).detach(), HM=torch.stack(hm, 0).detach(), dz=xw.grad, db=b.grad)


@pytest.mark.parametrize("T,K,N", [(40, 2, 77), (33, 2, 32), (5, 1, 1), (12, 2, 100)])
def test_tensor_core_sequence_kernels_match_float64_autograd_and_the_fma_kernels(T, K, N):
    torch, _lib, L, dev = _setup()
    d = _seq_inputs(torch, dev, T, K, N, seed=T + N)
    assert L.irrl_lstm_seq_set_path(-1) == 0, "the tensor-core recurrence is the default path"
    a, b = _run_seq(torch, _lib, L, dev, 0, d, T, K, N), _run_seq(torch, _lib, L, dev, 1, d, T, K, N)
    ref = _seq_reference(torch, d, T, K, N)
    for k in ("Hs", "Cs", "HM", "dz", "db"):
        scale = float(ref[k].abs().max()) + 1e-30
        for name, r in (("tensor-core", a), ("fma", b)):
            err = float((r[k].double() - ref[k]).abs().max())
            assert err <= 3e-6 * scale, (name, k, err, scale)
    for k in a:                                           # the two kernel families agree with each other at the same level
        assert float((a[k] - b[k]).abs().max()) <= 4e-6 * (float(b[k].abs().max()) + 1e-30), k


@pytest.mark.parametrize("T,K,N", [(7, 2, 37), (5, 2, 77), (3, 2, 64), (4, 1, 200), (2, 2, 1), (9, 2, 31)])
def test_streaming_products_match_float64(T, K, N):
    """irrl_proj_rows (x W_x for the shared observation and for a per-tower input, dz W_x^T) and irrl_gram_rows (dW = sum x^T dz)
    on ragged tiles (N not a multiple of the 64- / 32-row tiles, N = 1), through the autograd wrapper the learner uses"""
    torch, _lib, L, dev = _setup()
    from high_speed_quadrupedal_locomotion_by_irrl_b200.lstm_seq import ProjRows
    g = torch.Generator(device=dev); g.manual_seed(N)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    obs, H0, D = r(T, N, 35), r(T, K, N, 48), r(T, K, N, 192)
    wx0, wx1 = r(K, 35, 192) * 0.3, r(K, 48, 192) * 0.3
    for X, W in ((obs, wx0), (H0, wx1)):
        Xr = X.clone().requires_grad_(X.dim() == 4); Wr = W.clone().requires_grad_(True)
        Y = ProjRows.apply(Xr, Wr); (Y * D).sum().backward()
        Xd = X.double(); Xk = Xd.unsqueeze(1).expand(T, K, N, -1) if X.dim() == 3 else Xd
        Yd = torch.matmul(Xk, W.double()); dW = torch.matmul(Xk.transpose(-1, -2), D.double()).sum(0)
        assert float((Y.detach().double() - Yd).abs().max()) <= 3e-6 * float(Yd.abs().max())
        assert float((Wr.grad.double() - dW).abs().max()) <= 3e-6 * float(dW.abs().max())
        if X.dim() == 4:
            dX = torch.matmul(D.double(), W.double().transpose(1, 2))
            assert float((Xr.grad.double() - dX).abs().max()) <= 3e-6 * float(dX.abs().max())


def test_streaming_products_reject_unsupported_shapes():
    torch, _lib, L, dev = _setup()
    x = torch.zeros(2, 2, 8, 64, device=dev); w = torch.zeros(2, 64, 192, device=dev); y = torch.zeros(2, 2, 8, 192, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    assert L.irrl_proj_rows(None, 2, 2, 8, p(x), 64, 1, p(w), 0, p(y), 192) != 0
    assert b"unsupported shape" in L.irrl_last_error()
    assert L.irrl_gram_rows(None, 2, 2, 8, p(x), 64, 1, p(y), p(y)) != 0


def test_learner_gradients_do_not_depend_on_the_kernel_family():
    """one PPO loss / gradient evaluation at a rollout-like size: tensor-core kernels + the fused heads / loss kernel (default) against the
    FMA recurrence with torch.matmul projections and the autograd loss (IRRL_LEARNER_GEMM=cublas, path 1) -- loss, statistics and all 17
    used parameter gradients (ent_coef != 0 so that the logstd gradient has both its parts)"""
    import os
    torch, _lib, L, dev = _setup()
    from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import LstmActorCritic, ppo_loss
    from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bp5_155_params.npz")); W = [z[k] for k in PARAM_NAMES]
    rng = np.random.default_rng(5); T, N = 64, 200
    t = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
    obs, masks, st = t(rng.normal(0, 0.7, (T, N, 35))), t(rng.random((T, N)) < 0.05), t(rng.normal(size=(N, 384)) * 0.3)
    act, adv, ret = t(rng.normal(0, 0.3, (T, N, 12))), t(rng.normal(size=(T, N))), t(rng.normal(size=(T, N)))
    oldv, oldn = t(rng.normal(size=(T, N))), t(rng.normal(-15, 1, (T, N)))
    res = []
    for path, gemm in ((0, ""), (1, "cublas")):
        prev = L.irrl_lstm_seq_set_path(path); os.environ["IRRL_LEARNER_GEMM"] = gemm
        try:
            m = LstmActorCritic(W).to(dev)
            loss, stats = ppo_loss(m, obs, masks, st, act, adv, ret, oldv, oldn, 0.2, 0.01, 0.5, time_major=True)
            res.append((float(loss.detach()), torch.autograd.grad(loss, m.param_list(), allow_unused=True), {k: float(v) for k, v in stats.items()}))
        finally:
            L.irrl_lstm_seq_set_path(prev); os.environ.pop("IRRL_LEARNER_GEMM", None)
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * max(1.0, abs(res[1][0]))
    for k in ("policy_loss", "value_loss", "policy_entropy", "approxkl", "clipfrac"):          # the fused heads + loss kernel reports the same statistics
        assert abs(res[0][2][k] - res[1][2][k]) <= 1e-5 * max(1.0, abs(res[1][2][k])), (k, res[0][2][k], res[1][2][k])
    for n, ga, gb in zip(PARAM_NAMES, res[0][1], res[1][1]):
        if gb is None:
            assert ga is None; continue
        assert float((ga - gb).abs().max()) <= 2e-4 * (float(gb.abs().max()) + 1e-12), n


@pytest.mark.parametrize("T,K,N,cx", [(7, 2, 37, 35), (5, 2, 76, 48), (3, 2, 64, 48), (4, 1, 200, 35), (2, 2, 4, 48), (9, 2, 44, 35), (6, 2, 24, 48), (3, 2, 1, 48)])
def test_tcgen05_weight_gradients_match_float64(T, K, N, cx):
    """irrl_gram2_rows: dW_x and dW_h of a layer from one pass over dz, accumulators in tensor memory over all tiles of a CTA; ragged
    32-row tiles, the shared 35-column observation and a per-tower 48-column input"""
    torch, _lib, L, dev = _setup()
    from high_speed_quadrupedal_locomotion_by_irrl_b200.lstm_seq import gram2_rows
    g = torch.Generator(device=dev); g.manual_seed(N + cx)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    from high_speed_quadrupedal_locomotion_by_irrl_b200.lstm_seq import masked_input_state
    X = r(T, N, cx) if cx == 35 else r(T, K, N, cx)
    Hs, h0, D = r(T, K, N, 48), r(K, N, 48), r(T, K, N, 192)
    keep = (torch.rand(T, N, device=dev, generator=g) > 0.2).float()
    HM = masked_input_state(Hs, h0, keep)               # hm(t) = Hs(t-1) * keep(t): the kernel forms it on the fly
    dwx, dwh = gram2_rows(X, Hs, h0, keep, D)
    Xk = (X.double().unsqueeze(1).expand(T, K, N, -1) if X.dim() == 3 else X.double())
    rx = torch.matmul(Xk.transpose(-1, -2), D.double()).sum(0); rh = torch.matmul(HM.double().transpose(-1, -2), D.double()).sum(0)
    assert dwx.shape == rx.shape and dwh.shape == rh.shape
    assert float((dwx.double() - rx).abs().max()) <= 3e-6 * float(rx.abs().max())
    assert float((dwh.double() - rh).abs().max()) <= 3e-6 * float(rh.abs().max())


def test_forward_projection_kernels_agree():
    """the tcgen05 projection kernel (default) and the warp-level MMA kernel compute the same 3xTF32 products"""
    torch, _lib, L, dev = _setup()
    p = lambda t: C.c_void_p(t.data_ptr())
    g = torch.Generator(device=dev); g.manual_seed(3)
    T, K, N = 6, 2, 300
    for X, W, c, tw in ((torch.randn(T, N, 35, device=dev, generator=g), torch.randn(K, 35, 192, device=dev, generator=g), 35, 0),
                        (torch.randn(T, K, N, 48, device=dev, generator=g), torch.randn(K, 48, 192, device=dev, generator=g), 48, 1)):
        out = []
        for path in (0, 1):
            prev = L.irrl_proj_rows_set_path(path)
            try:
                Y = torch.full((T, K, N, 192), float("nan"), device=dev)
                _lib.check(L.irrl_proj_rows(None, T, K, N, p(X), c, tw, p(W), 0, p(Y), 192)); torch.cuda.synchronize()
            finally:
                L.irrl_proj_rows_set_path(prev)
            out.append(Y)
        ref = torch.matmul(X.double().unsqueeze(1) if tw == 0 else X.double(), W.double())
        for Y in out:
            assert float((Y.double() - ref).abs().max()) <= 3e-6 * float(ref.abs().max())


def test_fused_head_loss_scales_with_the_upstream_gradient():
    """the activation gradient of the fused heads + loss kernel is scaled on the device only when the upstream gradient is not 1"""
    torch, _lib, L, dev = _setup()
    from high_speed_quadrupedal_locomotion_by_irrl_b200.lstm_seq import HeadLossFused
    g = torch.Generator(device=dev); g.manual_seed(11)
    r = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
    T, N = 5, 52
    args = (r(48, 12, sc=0.1), r(12, sc=0.1), r(48, 1, sc=0.1), r(1), r(1, 12, sc=0.2), r(T, N, 12), r(T, N), r(T, N), r(T, N), r(T, N) - 15.0)
    out = []
    for scale in (1.0, 2.5):
        H1 = r(T, 2, N, 48).requires_grad_(True) if not out else out[0][2].detach().clone().requires_grad_(True)
        pw = args[0].clone().requires_grad_(True)
        loss, _ = HeadLossFused.apply(H1, pw, *args[1:], 0.2, 0.5)
        gh, gw = torch.autograd.grad(loss * scale, (H1, pw))
        out.append((gh, gw, H1))
    assert torch.allclose(out[1][0], 2.5 * out[0][0], rtol=1e-6, atol=0) and torch.allclose(out[1][1], 2.5 * out[0][1], rtol=1e-5, atol=1e-12)


def test_streaming_kernels_with_many_tiles_per_cta():
    """the persistent tcgen05 kernels in their steady state: several tiles per CTA (stage / slot / TMEM-buffer reuse, barrier parities wrapping),
    ragged last tiles, against float64"""
    torch, _lib, L, dev = _setup()
    from high_speed_quadrupedal_locomotion_by_irrl_b200.lstm_seq import ProjRows, gram2_rows, masked_input_state
    g = torch.Generator(device=dev); g.manual_seed(21)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    T, K, N = 160, 2, 520                    # 800 projection tiles and 3520 weight-gradient tiles per tower on 74 CTAs
    obs, H0, D = r(T, N, 35), r(T, K, N, 48), r(T, K, N, 192)
    wx0, wx1 = r(K, 35, 192) * 0.3, r(K, 48, 192) * 0.3
    for X, W in ((obs, wx0), (H0, wx1)):
        Xr = X.clone().requires_grad_(X.dim() == 4); Wr = W.clone().requires_grad_(True)
        Y = ProjRows.apply(Xr, Wr); (Y * D).sum().backward()
        Xk = X.double().unsqueeze(1).expand(T, K, N, -1) if X.dim() == 3 else X.double()
        Yd = torch.matmul(Xk, W.double())
        assert float((Y.detach().double() - Yd).abs().max()) <= 3e-6 * float(Yd.abs().max())
        if X.dim() == 4:
            dX = torch.matmul(D.double(), W.double().transpose(1, 2))
            assert float((Xr.grad.double() - dX).abs().max()) <= 3e-6 * float(dX.abs().max())
        h0s = r(K, N, 48); keep = (torch.rand(T, N, device=dev, generator=g) > 0.2).float()
        dwx, dwh = gram2_rows(X, H0, h0s, keep, D)
        HMd = masked_input_state(H0, h0s, keep).double()
        rx = torch.matmul(Xk.transpose(-1, -2), D.double()).sum(0); rh = torch.matmul(HMd.transpose(-1, -2), D.double()).sum(0)
        # 83 200-row reductions of O(1) random products accumulated in fp32 (tensor memory, then 74 partials): the bar is the rounding noise of an
        # fp32 accumulator that wanders to ~1e3 (measured 8e-6 of the scale; torch.matmul in fp32 sits at the same level)
        assert float((dwx.double() - rx).abs().max()) <= 2e-5 * float(rx.abs().max())
        assert float((dwh.double() - rh).abs().max()) <= 2e-5 * float(rh.abs().max())
