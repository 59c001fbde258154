"""CPU tests of the learner side: torch policy vs the numpy oracle, PPO loss vs ppo2.py:152-175 restated
in numpy, TF1 Adam, global-norm clipping, and the world_size-2 gloo path (advantage statistics + gradient all-reduce)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import PARAM_NAMES, init_params, flatten_params, unflatten_params
from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import LstmActorCritic, TF1Adam, clip_by_global_norm, ppo_loss, global_mean_std, allreduce_grads
from oracle import lstm_oracle as LO

G = os.path.join(os.path.dirname(__file__), "golden")


def _weights():
    z = np.load(os.path.join(G, "bp5_155_params.npz"))
    return [z[k] for k in PARAM_NAMES]


def _batch(rng, N, T):
    obs = rng.normal(size=(N, T, 35)).astype(np.float32); masks = (rng.random((N, T)) < 0.15).astype(np.float32)
    st = (rng.normal(size=(N, 384)) * 0.3).astype(np.float32); act = rng.normal(0, 0.3, size=(N, T, 12)).astype(np.float32)
    adv = rng.normal(size=(N, T)).astype(np.float32); ret = rng.normal(size=(N, T)).astype(np.float32)
    oldv = rng.normal(size=(N, T)).astype(np.float32); oldn = rng.normal(-15, 1, size=(N, T)).astype(np.float32)
    return obs, masks, st, act, adv, ret, oldv, oldn


def test_torch_policy_matches_numpy_oracle_over_masked_sequence():
    W = _weights(); m = LstmActorCritic(W)
    rng = np.random.default_rng(0)
    obs, masks, st, *_ = _batch(rng, 5, 9)
    mean, val, ns = m.forward_sequence(torch.tensor(obs), torch.tensor(masks), torch.tensor(st))
    s = st.astype(np.float64)
    for t in range(obs.shape[1]):
        a, v, s, nlp, mu = LO.act(dict(zip(PARAM_NAMES, W)), obs[:, t], s, masks[:, t])
        assert np.abs(mu - mean[:, t].detach().numpy()).max() < 5e-5 and np.abs(v - val[:, t].detach().numpy()).max() < 5e-5   # fp32 torch vs fp64 numpy
    assert np.abs(ns.detach().numpy() - s).max() < 5e-5


def test_param_flatten_roundtrip_and_count():
    W = init_params(np.random.default_rng(1))
    flat = flatten_params(W)
    assert flat.size == 70741
    for a, b in zip(W, unflatten_params(flat)):
        assert np.array_equal(a, b)


def test_ppo_loss_matches_numpy_restatement():
    W = _weights(); m = LstmActorCritic(W)
    rng = np.random.default_rng(2)
    obs, masks, st, act, adv, ret, oldv, oldn = _batch(rng, 4, 6)
    t = lambda x: torch.tensor(x)
    loss, stats = ppo_loss(m, t(obs), t(masks), t(st), t(act), t(adv), t(ret), t(oldv), t(oldn), 0.2, 0.0, 0.5)
    # numpy restatement of ppo2.py:152-175
    P = dict(zip(PARAM_NAMES, W)); s = st.astype(np.float64); nlp = np.zeros((4, 6)); vp = np.zeros((4, 6))
    for k in range(6):
        a, v, s, _, mu = LO.act(P, obs[:, k], s, masks[:, k])
        std = np.exp(P["pi_logstd"].reshape(1, -1))
        nlp[:, k] = 0.5 * (((act[:, k] - mu) / std) ** 2).sum(1) + 0.5 * np.log(2 * np.pi) * 12 + P["pi_logstd"].sum()
        vp[:, k] = v
    vclip = oldv + np.clip(vp - oldv, -0.2, 0.2)
    vf = 0.5 * np.maximum((vp - ret) ** 2, (vclip - ret) ** 2).mean()
    ratio = np.exp(oldn - nlp)
    pg = np.maximum(-adv * ratio, -adv * np.clip(ratio, 0.8, 1.2)).mean()
    assert abs(float(loss) - (pg + 0.5 * vf)) < 1e-4 * max(1.0, abs(pg + 0.5 * vf))
    assert abs(float(stats["value_loss"]) - vf) < 1e-4 * max(1.0, vf)
    # gradients exist for everything except the unused q head
    grads = torch.autograd.grad(loss, m.param_list(), allow_unused=True)
    for n, g in zip(PARAM_NAMES, grads):
        assert (g is None) == n.startswith("q_"), n


def test_tf1_adam_and_global_norm_clip():
    p = [torch.tensor([1.0, -2.0]), torch.tensor([[0.5]])]
    g = [torch.tensor([0.3, -0.1]), torch.tensor([[2.0]])]
    clipped, gn = clip_by_global_norm(g, 0.5)
    ref_norm = np.sqrt(0.3 ** 2 + 0.1 ** 2 + 2.0 ** 2)
    assert abs(float(gn) - ref_norm) < 1e-6
    assert abs(float(torch.sqrt(sum((c ** 2).sum() for c in clipped))) - 0.5) < 1e-6
    small, _ = clip_by_global_norm([torch.tensor([0.1])], 0.5)
    assert float(small[0]) == pytest.approx(0.1)                    # norms below the threshold are untouched
    opt = TF1Adam(p, eps=1e-5)
    before = [x.clone() for x in p]
    opt.step(g, lr=1e-3)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    for b, x, gg in zip(before, p, g):
        m = 0.1 * gg; v = 0.001 * gg * gg
        assert torch.allclose(x, b - lr_t * m / (torch.sqrt(v) + 1e-5), atol=1e-9)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    rng = np.random.default_rng(3)
    W = init_params(np.random.default_rng(4))
    m = LstmActorCritic(W)
    obs, masks, st, act, adv, ret, oldv, oldn = _batch(rng, 6, 5)       # the global batch: 6 envs
    lo, hi = rank * 3, rank * 3 + 3                                     # contiguous env shard
    t = lambda x: torch.tensor(x[lo:hi])
    a = t(ret) - t(oldv)
    mean, std = global_mean_std(a)
    full = torch.tensor(ret - oldv)
    ok_stats = bool(abs(float(mean) - float(full.mean())) < 1e-6 and abs(float(std) - float(full.std(unbiased=False))) < 1e-6)
    an = (a - mean) / (std + 1e-8)
    loss, _ = ppo_loss(m, t(obs), t(masks), t(st), t(act), an, t(ret), t(oldv), t(oldn), 0.2, 0.0, 0.5)
    grads = torch.autograd.grad(loss, m.param_list(), allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, m.param_list())]
    grads = allreduce_grads(grads)
    # single-process gradient on the concatenated batch
    fa = (full - full.mean()) / (full.std(unbiased=False) + 1e-8)
    tl = lambda x: torch.tensor(x)
    loss_full, _ = ppo_loss(m, tl(obs), tl(masks), tl(st), tl(act), fa, tl(ret), tl(oldv), tl(oldn), 0.2, 0.0, 0.5)
    gfull = torch.autograd.grad(loss_full, m.param_list(), allow_unused=True)
    gfull = [g if g is not None else torch.zeros_like(p) for g, p in zip(gfull, m.param_list())]
    err = max(float((a_ - b_).abs().max()) for a_, b_ in zip(grads, gfull))
    q.put((rank, ok_stats, err))
    dist.destroy_process_group()


def test_world_size_2_gloo_advantage_stats_and_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_stats, err in res:
        assert ok_stats, rank
        assert err < 2e-6, (rank, err)      # mean over equal shards == mean over the concatenated batch


def test_time_major_forward_equals_env_major_forward_cpu():
    W = _weights(); m = LstmActorCritic(W)
    rng = np.random.default_rng(6)
    obs, masks, st, *_ = _batch(rng, 5, 8)
    mean_a, val_a, _ = m.forward_sequence(torch.tensor(obs), torch.tensor(masks), torch.tensor(st))
    mean_b, val_b = m.forward_time_major(torch.tensor(obs).transpose(0, 1).contiguous(), 1.0 - torch.tensor(masks).t().contiguous(), torch.tensor(st), fused=False)
    assert torch.allclose(mean_a.transpose(0, 1), mean_b, atol=1e-5) and torch.allclose(val_a.t(), val_b, atol=1e-5)


def test_learner_kernel_wrappers_fall_back_off_the_gpu():
    """lstm_seq's tensor-core wrappers are CUDA-only: on CPU tensors the learner takes the plain torch path (no CUDA library is touched),
    and the time-major forward agrees with the env-major reference forward"""
    import torch
    from high_speed_quadrupedal_locomotion_by_irrl_b200 import lstm_seq
    from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import LstmActorCritic
    g = torch.Generator().manual_seed(3)
    X = torch.randn(4, 5, 35, generator=g); W = torch.randn(2, 35, 192, generator=g)
    assert not lstm_seq.fused_layer_ok(X, W)
    assert torch.equal(lstm_seq.proj_rows(X, W), torch.matmul(X.unsqueeze(1), W))
    m = LstmActorCritic()
    T, N = 6, 5
    obs = torch.randn(T, N, 35, generator=g); masks = (torch.rand(T, N, generator=g) < 0.2).float(); st = torch.randn(N, 384, generator=g) * 0.3
    mean_t, val_t = m.forward_time_major(obs, 1.0 - masks, st)
    mean_e, val_e, _ = m.forward_sequence(obs.transpose(0, 1), masks.transpose(0, 1), st)
    assert torch.allclose(mean_t, mean_e.transpose(0, 1), atol=1e-5) and torch.allclose(val_t, val_e.transpose(0, 1), atol=1e-5)
    H1, own = m.features_time_major(obs, 1.0 - masks, st)
    assert own is False and H1.shape == (T, 2, N, 48)


def test_closed_form_head_loss_gradients_match_autograd():
    """the per-sample gradient formulas the fused heads + loss kernel evaluates (csrc/ppo_head_loss.cu: d loss / d mean, d loss / d v, d loss / d logstd
    with the sub-gradient conventions of torch.maximum / torch.clamp), restated in torch and checked against autograd of ppo_loss's arithmetic"""
    import math
    import torch
    g = torch.Generator().manual_seed(7)
    B, eps, vf_coef = 4000, 0.2, 0.5
    mean = torch.randn(B, 12, generator=g, dtype=torch.float64).requires_grad_(True)
    v = torch.randn(B, generator=g, dtype=torch.float64).requires_grad_(True)
    logstd = (torch.randn(1, 12, generator=g, dtype=torch.float64) * 0.3).requires_grad_(True)
    act = mean.detach() + torch.randn(B, 12, generator=g, dtype=torch.float64) * 0.5
    adv = torch.randn(B, generator=g, dtype=torch.float64); ret = torch.randn(B, generator=g, dtype=torch.float64)
    oldv = v.detach() + torch.randn(B, generator=g, dtype=torch.float64) * 0.3
    with torch.no_grad():
        nlp0 = 0.5 * (((act - mean) / logstd.exp()) ** 2).sum(-1) + 0.5 * math.log(2 * math.pi) * 12 + logstd.sum()
    oldn = nlp0 + torch.randn(B, generator=g, dtype=torch.float64) * 0.3          # ratios on both sides of the clip range
    # autograd (the arithmetic of ppo2.ppo_loss)
    nlp = 0.5 * (((act - mean) / logstd.exp()) ** 2).sum(-1) + 0.5 * math.log(2 * math.pi) * 12 + logstd.sum()
    ratio = torch.exp(oldn - nlp)
    pg = torch.maximum(-adv * ratio, -adv * torch.clamp(ratio, 1 - eps, 1 + eps)).mean()
    vclip = oldv + torch.clamp(v - oldv, -eps, eps)
    vf = 0.5 * torch.maximum((v - ret) ** 2, (vclip - ret) ** 2).mean()
    gm, gv, gl = torch.autograd.grad(pg + vf_coef * vf, (mean, v, logstd))
    # closed form, as in the kernel
    with torch.no_grad():
        istd = torch.exp(-logstd); z = (act - mean) * istd
        r = ratio.detach(); rc = r.clamp(1 - eps, 1 + eps); pg1, pg2 = -adv * r, -adv * rc
        inside = ((r >= 1 - eps) & (r <= 1 + eps)).double()
        sel = torch.where(pg1 > pg2, torch.ones_like(r), torch.where(pg2 > pg1, inside, 0.5 + 0.5 * inside))
        g_nlp = adv * r * sel / B
        dv = v - oldv; e1 = v - ret; e2 = oldv + dv.clamp(-eps, eps) - ret; l1, l2 = e1 * e1, e2 * e2
        inside_v = ((dv >= -eps) & (dv <= eps)).double()
        dvf = torch.where(l1 > l2, 2 * e1, torch.where(l2 > l1, 2 * e2 * inside_v, e1 + e2 * inside_v))
        cm = -g_nlp.unsqueeze(1) * z * istd; cv = vf_coef * 0.5 * dvf / B; cl = (g_nlp.unsqueeze(1) * (1 - z * z)).sum(0, keepdim=True)
    assert (0.05 < float((sel == 0).double().mean()) < 0.95) and (0.05 < float((l2 > l1).double().mean()) < 0.95)      # both branches of both clips are exercised
    assert torch.allclose(cm, gm, rtol=1e-10, atol=1e-14) and torch.allclose(cv, gv, rtol=1e-10, atol=1e-14) and torch.allclose(cl, gl, rtol=1e-9, atol=1e-13)
