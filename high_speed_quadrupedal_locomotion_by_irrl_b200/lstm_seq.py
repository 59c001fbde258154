"""Manual BPTT through one LSTM layer of BOTH towers over a whole rollout, for the PPO learner (ppo2.py:136-197).

LstmLayerSeqPersistent: the whole time loop inside one kernel per direction (irrl_lstm_seq_fwd / _bwd, recurrent products on the
tensor cores); ProjRows: the input projections x W_x of all T steps, their input gradient and the weight gradients
dW = sum_t x_t^T dz_t as streaming tensor-core kernels (irrl_proj_rows / irrl_gram_rows) -- IRRL_LEARNER_GEMM=cublas keeps
torch.matmul for A/B.  LstmLayerSeq: the older step-wise path [batched h W_h GEMM (cuBLAS) -> fused cell kernel], two launches per
step.  Layout is time-major ([T, tower, env, ...]) -- exactly how the device rollout stores mb_* -- so every step is a contiguous slice.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _own_gemm() -> bool:
    return os.environ.get("IRRL_LEARNER_GEMM", "") != "cublas"


def gram_rows(X, D):
    """sum over (t, n) of X[t,(k),n,:]^T D[t,k,n,:] -> [K, x_cols, 192]; X [T,N,c] (shared by the towers) or [T,K,N,c], c <= 48."""
    L = _lib.load()
    T, K, N, _ = D.shape
    c = X.shape[-1]
    st = C.c_void_p(torch.cuda.current_stream(D.device).cuda_stream)
    part = D.new_empty((L.irrl_gram_rows_ctas(T, K, N), K, 48, 192))
    _lib.check(L.irrl_gram_rows(st, T, K, N, _p(X), c, 1 if X.dim() == 4 else 0, _p(D), _p(part)), "gram_rows")
    return part.sum(0)[:, :c]


def masked_input_state(Hs, h0, keep):
    """HM[t] = h(t-1) * keep[t], the state SB's lstm() feeds into step t (h0 * keep[0] at t = 0): [T,K,N,48]"""
    T, K, N, _ = Hs.shape
    prev = torch.cat([h0.unsqueeze(0), Hs[:-1]], 0)
    return prev * keep.view(T, 1, N, 1)


def gram2_rows(X, Hs, h0, keep, D):
    """dW_x = sum x^T dz and dW_h = sum hm^T dz of one layer in one pass over dz (tcgen05), hm(t) = Hs(t-1) * keep(t) formed on the fly from
    the layer's output sequence: ([K,c,192], [K,48,192])."""
    L = _lib.load()
    T, K, N, _ = D.shape
    c = X.shape[-1]
    st = C.c_void_p(torch.cuda.current_stream(D.device).cuda_stream)
    Hs = Hs.contiguous(); h0 = h0.contiguous(); keep = keep.contiguous()
    if N % 4 or (X.data_ptr() | Hs.data_ptr() | h0.data_ptr() | keep.data_ptr() | D.data_ptr()) % 16:
        # the tcgen05 kernel streams whole 16-byte aligned blocks: odd batches go through the warp-level kernel
        return gram_rows(X, D), gram_rows(masked_input_state(Hs, h0, keep), D)
    part = D.new_empty((L.irrl_gram2_rows_ctas(T, K, N), K, 128, 192))
    _lib.check(L.irrl_gram2_rows(st, T, K, N, _p(X), c, 1 if X.dim() == 4 else 0, _p(Hs), _p(h0), _p(keep), _p(D), _p(part)), "gram2_rows")
    G = part[:, :, :96].sum(0)
    return G[:, :c], G[:, 48:96]


class LstmLayerFused(torch.autograd.Function):
    """One LSTM layer of both towers over the whole rollout, projection included: H = lstm(X W_x + b, W_h).  Forward = irrl_proj_rows
    (tcgen05) + irrl_lstm_seq_fwd; backward = irrl_lstm_seq_bwd, then BOTH weight gradients from one pass over dz (irrl_gram2_rows, tcgen05)
    and, for a layer fed by another layer, dX = dz W_x^T (irrl_proj_rows).  The projection xw is not kept for the backward pass."""

    @staticmethod
    def forward(ctx, X, wx, wh, b, c0, h0, keep):
        L = _lib.load()
        X = X.contiguous(); wx = wx.contiguous(); wh = wh.contiguous(); b = b.contiguous(); c0 = c0.contiguous(); h0 = h0.contiguous(); keep = keep.contiguous()
        K, c, _ = wx.shape
        T, N = X.shape[0], X.shape[-2]
        st = C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)
        xw = X.new_empty((T, K, N, 192))
        _lib.check(L.irrl_proj_rows(st, T, K, N, _p(X), c, 1 if X.dim() == 4 else 0, _p(wx), 0, _p(xw), 192), "proj_rows")
        gates = torch.empty_like(xw); Cs = X.new_empty((T, K, N, 48)); Hs = X.new_empty((T, K, N, 48))
        # no masked copy of the hidden state: the weight-gradient kernel forms hm(t) = Hs(t-1) * keep(t) itself
        _lib.check(L.irrl_lstm_seq_fwd(st, T, K, N, _p(xw), _p(wh), _p(c0), _p(h0), _p(keep), _p(gates), _p(Cs), _p(Hs), _p(b), None), "lstm_seq_fwd")
        ctx.save_for_backward(X, wx, wh, keep, gates, Cs, Hs, c0, h0)
        return Hs

    @staticmethod
    def backward(ctx, dH):
        L = _lib.load()
        X, wx, wh, keep, gates, Cs, Hs, c0, h0 = ctx.saved_tensors
        T, K, N, _ = gates.shape
        c = X.shape[-1]
        st = C.c_void_p(torch.cuda.current_stream(gates.device).cuda_stream)
        dH = dH.contiguous()
        DZ = torch.empty_like(gates)
        db_part = gates.new_empty((L.irrl_lstm_seq_ctas(N), K, 192))
        _lib.check(L.irrl_lstm_seq_bwd(st, T, K, N, _p(dH), _p(wh), _p(c0), _p(keep), _p(gates), _p(Cs), _p(DZ), _p(db_part)), "lstm_seq_bwd")
        dwx, dwh = gram2_rows(X, Hs, h0, keep, DZ)
        dX = None
        if ctx.needs_input_grad[0]:
            if X.dim() == 4 and c == 48:
                dX = torch.empty_like(X)
                _lib.check(L.irrl_proj_rows(st, T, K, N, _p(DZ), 192, 1, _p(wx), 1, _p(dX), c), "proj_rows^T")
            else:
                dX = torch.matmul(DZ, wx.transpose(1, 2)); dX = dX.sum(1) if X.dim() == 3 else dX
        return dX, dwx, dwh, db_part.sum(0), None, None, None


class HeadLossFused(torch.autograd.Function):
    """Heads + PPO2 loss (ppo2.py:152-175) and their gradients in one kernel (irrl_ppo_head_loss): returns
    (mean(pg) + vf_coef * 0.5 * mean(vf), [pg_loss, vf_loss, approxkl, clipfrac]); the entropy term stays outside."""

    @staticmethod
    def forward(ctx, H1, pi_w, pi_b, vf_w, vf_b, logstd, actions, advs, returns, old_values, old_neglogp, cliprange, vf_coef):
        L = _lib.load()
        T, K, N, _ = H1.shape
        H1 = H1.contiguous()
        c = lambda x: x.detach().contiguous()
        st = C.c_void_p(torch.cuda.current_stream(H1.device).cuda_stream)
        dH = torch.empty_like(H1); G = H1.new_empty((T, N, 16))
        part = H1.new_empty((L.irrl_ppo_head_loss_ctas(T, N), 16))
        inv = 1.0 / float(T * N)
        a = [c(pi_w), c(pi_b), c(vf_w), c(vf_b), c(logstd), c(actions), c(advs), c(returns), c(old_values), c(old_neglogp)]
        _lib.check(L.irrl_ppo_head_loss(st, T, N, _p(H1), *[_p(x) for x in a], float(cliprange), float(vf_coef), inv, _p(dH), _p(G), _p(part)), "ppo_head_loss")
        s = part.sum(0)
        pg, vf = s[0] * inv, 0.5 * s[1] * inv
        stats = torch.stack([pg, vf, s[2] * inv, s[3] * inv])
        ctx.save_for_backward(H1, G, dH, s[4:16].clone())
        ctx.mark_non_differentiable(stats)
        return pg + float(vf_coef) * vf, stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        H1, G, dH, dls = ctx.saved_tensors
        gm = G[..., :12]
        d_pi_w = torch.matmul(H1[:, 0].transpose(-1, -2), gm).sum(0)               # T batched [48 x N] x [N x 12] products on strided views
        d_vf_w = torch.matmul(H1[:, 1].transpose(-1, -2), G[..., 12:13]).sum(0)
        d_pi_b = gm.sum((0, 1)); d_vf_b = G[..., 12].sum().reshape(1)
        g = g_loss
        # the 2.4 GB activation gradient is scaled in place, and only if the upstream gradient is not 1 (it is 1 for a terminal loss)
        L = _lib.load()
        gs = g.detach().reshape(1).float().contiguous()
        _lib.check(L.irrl_scale_unless_one(C.c_void_p(torch.cuda.current_stream(dH.device).cuda_stream), _p(dH), dH.numel(), _p(gs)), "scale")
        return (dH, d_pi_w * g, d_pi_b * g, d_vf_w * g, d_vf_b * g, dls.reshape(1, 12) * g, None, None, None, None, None, None, None)


def fused_layer_ok(X, wx) -> bool:
    K, c, n_out = wx.shape
    return X.is_cuda and _own_gemm() and os.environ.get("IRRL_LEARNER_FUSED", "1") != "0" and n_out == 192 and (c == 48 or (32 < c <= 40 and X.dim() == 3))


class ProjRows(torch.autograd.Function):
    """Y[T,K,N,192] = X . W_k for all T steps; X [T,N,c] (the observation, shared by both towers) or [T,K,N,48]; W [K,c,192]."""

    @staticmethod
    def forward(ctx, X, W):
        L = _lib.load()
        X = X.contiguous(); W = W.contiguous()
        K, c, n_out = W.shape
        T, N = X.shape[0], X.shape[-2]
        st = C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)
        Y = X.new_empty((T, K, N, n_out))
        _lib.check(L.irrl_proj_rows(st, T, K, N, _p(X), c, 1 if X.dim() == 4 else 0, _p(W), 0, _p(Y), n_out), "proj_rows")
        ctx.save_for_backward(X, W)
        return Y

    @staticmethod
    def backward(ctx, dY):
        L = _lib.load()
        X, W = ctx.saved_tensors
        dY = dY.contiguous()
        T, K, N, n_out = dY.shape
        c = X.shape[-1]
        dW = gram_rows(X, dY) if ctx.needs_input_grad[1] else None
        dX = None
        if ctx.needs_input_grad[0]:
            if X.dim() == 4 and c == 48:
                st = C.c_void_p(torch.cuda.current_stream(dY.device).cuda_stream)
                dX = torch.empty_like(X)
                _lib.check(L.irrl_proj_rows(st, T, K, N, _p(dY), n_out, 1, _p(W), 1, _p(dX), c), "proj_rows^T")
            else:
                dX = torch.matmul(dY, W.transpose(1, 2))
                dX = dX.sum(1) if X.dim() == 3 else dX
        return dX, dW


def proj_rows(X, W):
    """x W_x for every step and tower: tensor-core streaming kernel on CUDA for the learner's shapes, torch.matmul otherwise."""
    K, c, n_out = W.shape
    if X.is_cuda and _own_gemm() and n_out == 192 and (c == 48 or (c <= 40 and X.dim() == 3)):
        return ProjRows.apply(X, W)
    return torch.matmul(X.unsqueeze(1) if X.dim() == 3 else X, W)


def _with_bias(fn):
    """(xw, wh, b, c0, h0, keep) front end for the paths that want the bias already added to the projection"""
    def run(xw, wh, b, c0, h0, keep):
        return fn(xw + b.view(1, b.shape[0], 1, -1), wh, c0, h0, keep)
    return run


class LstmLayerSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xw, wh, c0, h0, keep):
        """xw [T,K,N,192] (x W_x + b), wh [K,48,192], c0/h0 [K,N,48] state before step 0 (unmasked), keep [T,N] = 1 - mask.
        Returns H [T,K,N,48] (unmasked outputs), c_T, h_T [K,N,48]."""
        L = _lib.load()
        T, K, N, _ = xw.shape
        rows = K * N
        st = C.c_void_p(torch.cuda.current_stream(xw.device).cuda_stream)
        xw = xw.contiguous(); keep = keep.contiguous()
        gates = torch.empty_like(xw); Cs = xw.new_empty((T, K, N, 48)); Hs = xw.new_empty((T, K, N, 48))
        HM = xw.new_empty((T, K, N, 48)); CM = xw.new_empty((T, K, N, 48))          # masked h/c fed INTO step t
        HM[0] = h0 * keep[0].view(1, N, 1); CM[0] = c0 * keep[0].view(1, N, 1)
        hm_n = xw.new_empty((K, N, 48)); cm_n = xw.new_empty((K, N, 48))
        z = xw.new_empty((K, N, 192))
        for t in range(T):
            torch.baddbmm(xw[t], HM[t], wh, out=z)
            last = t + 1 == T
            _lib.check(L.irrl_lstm_pw_fwd(st, rows, N, _p(z), _p(CM[t]), _p(keep[t + 1]) if not last else None, _p(gates[t]), _p(Cs[t]), _p(Hs[t]),
                                          _p(HM[t + 1]) if not last else _p(hm_n), _p(CM[t + 1]) if not last else _p(cm_n)))
        ctx.save_for_backward(wh, keep, gates, Cs, HM, CM)
        return Hs, Cs[T - 1].clone(), Hs[T - 1].clone()

    @staticmethod
    def backward(ctx, dH, dcT, dhT):
        L = _lib.load()
        wh, keep, gates, Cs, HM, CM = ctx.saved_tensors
        T, K, N, _ = gates.shape
        rows = K * N
        st = C.c_void_p(torch.cuda.current_stream(gates.device).cuda_stream)
        dH = dH.contiguous()
        DZ = torch.empty_like(gates)
        whT = wh.transpose(1, 2).contiguous()
        carry_h = gates.new_empty((K, N, 48)); carry_c = gates.new_empty((K, N, 48)); carry_c2 = gates.new_empty((K, N, 48))
        dh_last = dH[T - 1] + (dhT if dhT is not None else 0)
        for t in range(T - 1, -1, -1):
            last = t + 1 == T
            dh_in = dh_last if last else dH[t]
            _lib.check(L.irrl_lstm_pw_bwd(st, rows, N, _p(dh_in.contiguous()), None if last else _p(carry_h), None if last else _p(carry_c),
                                          None if last else _p(keep[t + 1]), _p(gates[t]), _p(Cs[t]), _p(CM[t]), _p(DZ[t]), _p(carry_c2)))
            carry_c, carry_c2 = carry_c2, carry_c
            torch.bmm(DZ[t], whT, out=carry_h)            # gradient w.r.t. the masked h fed to step t
        # weight gradient: one batched GEMM over all (t, env) rows
        dwh = torch.bmm(HM.permute(1, 3, 0, 2).reshape(K, 48, T * N), DZ.permute(1, 0, 2, 3).reshape(K, T * N, 192))
        return DZ, dwh, None, None, None


class LstmLayerSeqPersistent(torch.autograd.Function):
    """Same contract as LstmLayerSeq, but the whole time loop runs inside ONE kernel per direction (irrl_lstm_seq_fwd / _bwd):
    W_h resident in shared memory, the recurrent state on chip, no per-step launches and no Python in the loop."""

    @staticmethod
    def forward(ctx, xw, wh, b, c0, h0, keep):
        """xw [T,K,N,192] = x W_x WITHOUT the bias (b [K,192] is added inside the kernel); the kernel also writes the masked hidden state
        fed into every step (HM), which the weight gradient needs -- no extra passes over the big tensors in PyTorch."""
        L = _lib.load()
        T, K, N, _ = xw.shape
        st = C.c_void_p(torch.cuda.current_stream(xw.device).cuda_stream)
        xw = xw.contiguous(); keep = keep.contiguous(); wh = wh.contiguous(); c0 = c0.contiguous(); h0 = h0.contiguous(); b = b.contiguous()
        gates = torch.empty_like(xw); Cs = xw.new_empty((T, K, N, 48)); Hs = xw.new_empty((T, K, N, 48)); HM = xw.new_empty((T, K, N, 48))
        _lib.check(L.irrl_lstm_seq_fwd(st, T, K, N, _p(xw), _p(wh), _p(c0), _p(h0), _p(keep), _p(gates), _p(Cs), _p(Hs), _p(b), _p(HM)))
        ctx.save_for_backward(wh, keep, gates, Cs, HM, c0)
        return Hs, Cs[T - 1].clone(), Hs[T - 1].clone()

    @staticmethod
    def backward(ctx, dH, dcT, dhT):
        L = _lib.load()
        wh, keep, gates, Cs, HM, c0 = ctx.saved_tensors
        T, K, N, _ = gates.shape
        st = C.c_void_p(torch.cuda.current_stream(gates.device).cuda_stream)
        dH = dH.contiguous()
        if dhT is not None:
            dH = dH.clone(); dH[T - 1] += dhT
        DZ = torch.empty_like(gates)
        db_part = gates.new_empty((L.irrl_lstm_seq_ctas(N), K, 192))
        _lib.check(L.irrl_lstm_seq_bwd(st, T, K, N, _p(dH), _p(wh), _p(c0), _p(keep), _p(gates), _p(Cs), _p(DZ), _p(db_part)))
        # dW_h[k] = sum_{t,n} HM[t,k,n,:]^T DZ[t,k,n,:]: T*K batched [48 x N] x [N x 192] products on strided views (no copies, and
        # far more parallel than one skinny GEMM with a 1.5 M-long reduction), then a sum over T
        dwh = gram_rows(HM, DZ) if _own_gemm() else torch.matmul(HM.transpose(-1, -2), DZ).sum(0)
        return DZ, dwh, db_part.sum(0), None, None, None


def lstm_layer_reference(xw, wh, b, c0, h0, keep):
    """The same computation with plain autograd ops (CPU tests / cross-check of the manual backward)."""
    T, K, N, _ = xw.shape
    xw = xw + b.view(1, K, 1, -1)
    c, h = c0, h0
    out = []
    for t in range(T):
        k = keep[t].view(1, N, 1)
        c = c * k; h = h * k
        z = xw[t] + torch.bmm(h, wh)
        i, f, o, g = z.chunk(4, dim=2)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out, 0), c, h
