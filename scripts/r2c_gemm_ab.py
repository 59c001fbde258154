"""The learner's streaming tensor-core products (irrl_proj_rows / irrl_gram_rows) against float64 torch on ragged shapes, then device
   time per launch at the learner's size against torch.matmul (cuBLAS fp32).   python scripts/r2c_gemm_ab.py [envs] [steps]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
from high_speed_quadrupedal_locomotion_by_irrl_b200.lstm_seq import ProjRows, gram_rows
L = _lib.load()
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def check(T, K, N, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    obs, H0, D = r(T, N, 35), r(T, K, N, 48), r(T, K, N, 192)
    wx0, wx1 = r(K, 35, 192) * 0.3, r(K, 48, 192) * 0.3
    out = {}
    for name, X, W in (("obs", obs, wx0), ("h", H0, wx1)):
        Xr = X.clone().requires_grad_(X.dim() == 4); Wr = W.clone().requires_grad_(True)
        Y = ProjRows.apply(Xr, Wr); (Y * D).sum().backward()
        Xd = X.double(); Yd = torch.matmul(Xd.unsqueeze(1) if X.dim() == 3 else Xd, W.double())
        dW = torch.matmul((Xd.unsqueeze(1).expand(T, K, N, -1) if X.dim() == 3 else Xd).transpose(-1, -2), D.double()).sum(0)
        out[name + " Y"] = (float((Y.double() - Yd).abs().max()), float(Yd.abs().max()))
        out[name + " dW"] = (float((Wr.grad.double() - dW).abs().max()), float(dW.abs().max()))
        if X.dim() == 4:
            dX = torch.matmul(D.double(), W.double().transpose(1, 2))
            out[name + " dX"] = (float((Xr.grad.double() - dX).abs().max()), float(dX.abs().max()))
    print(f"T={T} K={K} N={N}: " + "  ".join(f"{k} {e:.2e}/{s:.1e}" for k, (e, s) in out.items()))
    assert all(e <= 3e-6 * s for e, s in out.values()), out


if __name__ == "__main__":
    for (T, K, N) in ((7, 2, 37), (5, 2, 77), (3, 2, 64), (4, 1, 200), (2, 2, 1), (9, 2, 31)):
        check(T, K, N, seed=N)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 750
    K = 2
    obs = torch.randn(T, N, 35, device=dev); H0 = torch.randn(T, K, N, 48, device=dev); D = torch.randn(T, K, N, 192, device=dev)
    wx0 = torch.randn(K, 35, 192, device=dev) * 0.3; wx1 = torch.randn(K, 48, 192, device=dev) * 0.3
    Y = torch.empty(T, K, N, 192, device=dev); dX = torch.empty(T, K, N, 48, device=dev)
    st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    gb = lambda *ts: sum(t.numel() for t in ts) * 4 / 1e9
    rows = [
        ("proj obs 35->192", lambda: L.irrl_proj_rows(st(), T, K, N, p(obs), 35, 0, p(wx0), 0, p(Y), 192), lambda: torch.matmul(obs.unsqueeze(1), wx0, out=Y), gb(obs, Y)),
        ("proj h   48->192", lambda: L.irrl_proj_rows(st(), T, K, N, p(H0), 48, 1, p(wx1), 0, p(Y), 192), lambda: torch.matmul(H0, wx1, out=Y), gb(H0, Y)),
        ("proj dz 192->48 ", lambda: L.irrl_proj_rows(st(), T, K, N, p(D), 192, 1, p(wx1), 1, p(dX), 48), lambda: torch.matmul(D, wx1.transpose(1, 2), out=dX), gb(D, dX)),
        ("gram h^T dz     ", lambda: gram_rows(H0, D), lambda: torch.matmul(H0.transpose(-1, -2), D).sum(0), gb(H0, D)),
        ("gram obs^T dz   ", lambda: gram_rows(obs, D), lambda: torch.matmul(obs.unsqueeze(1).transpose(-1, -2), D).sum(0), gb(obs, D)),
    ]
    for name, own, ref, g_ in rows:
        a, b = timed(own), timed(ref)
        extra = ""
        if name.startswith("proj") and "192->48" not in name:          # the forward projections default to the tcgen05 kernel: time the warp-level MMA kernel too
            L.irrl_proj_rows_set_path(1); c = timed(own); L.irrl_proj_rows_set_path(0)
            extra = f"   warp-level MMA kernel {c:6.2f} ms"
        print(f"{name} N={N} T={T}: tensor-core {a:6.2f} ms ({g_ / a * 1e3:5.0f} GB/s algorithmic)   torch.matmul fp32 {b:6.2f} ms{extra}")
    from high_speed_quadrupedal_locomotion_by_irrl_b200.lstm_seq import gram2_rows
    h0_ = torch.randn(K, N, 48, device=dev); keep_ = (torch.rand(T, N, device=dev) > 0.05).float()
    for name, X in (("gram2 [h | hm]^T dz  ", H0), ("gram2 [obs | hm]^T dz", obs)):
        a = timed(lambda: gram2_rows(X, H0, h0_, keep_, D))
        print(f"{name} N={N} T={T}: tcgen05 {a:6.2f} ms ({gb(X, H0, D) / a * 1e3:5.0f} GB/s algorithmic)   (two warp-level gram_rows launches above)")
