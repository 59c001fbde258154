"""Kernel-level A/B of the sequence-persistent BPTT kernels: tensor-core recurrence (path 0) against the FP32-FMA recurrence (path 1).
   python scripts/r2c_seq_ab.py [envs] [steps]      -> agreement on a ragged case, then device time per launch at the learner's size"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from high_speed_quadrupedal_locomotion_by_irrl_b200 import _lib
L = _lib.load()
dev = torch.device("cuda:0")
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def make(T, K, N, seed=0):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    r = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
    return dict(xw=r(T, K, N, 192), wh=r(K, 48, 192, sc=0.15), b=r(K, 192, sc=0.1), c0=r(K, N, 48, sc=0.5), h0=r(K, N, 48, sc=0.3),
                keep=(torch.rand(T, N, device=dev, generator=g) > 0.05).float(), dH=r(T, K, N, 48, sc=0.1))


def run(path, d, T, K, N):
    L.irrl_lstm_seq_set_path(path)
    gates = torch.empty(T, K, N, 192, device=dev); Cs = torch.empty(T, K, N, 48, device=dev); Hs = torch.empty_like(Cs); HM = torch.empty_like(Cs)
    _lib.check(L.irrl_lstm_seq_fwd(st(), T, K, N, p(d["xw"]), p(d["wh"]), p(d["c0"]), p(d["h0"]), p(d["keep"]), p(gates), p(Cs), p(Hs), p(d["b"]), p(HM)))
    dz = torch.empty_like(gates); db = torch.empty(L.irrl_lstm_seq_ctas(N), K, 192, device=dev)
    _lib.check(L.irrl_lstm_seq_bwd(st(), T, K, N, p(d["dH"]), p(d["wh"]), p(d["c0"]), p(d["keep"]), p(gates), p(Cs), p(dz), p(db)))
    torch.cuda.synchronize()
    return dict(gates=gates, Cs=Cs, Hs=Hs, HM=HM, dz=dz, db=db.sum(0))


def reference(d, T, K, N):
    """float64 autograd of the same recurrence"""
    xw = d["xw"].double().requires_grad_(True); wh = d["wh"].double(); b = d["b"].double()
    c, h = d["c0"].double(), d["h0"].double(); out = []
    for t in range(T):
        k = d["keep"][t].double().view(1, N, 1); c = c * k; h = h * k
        z = xw[t] + b.view(K, 1, 192) + torch.bmm(h, wh)
        i, f, o, g = z.chunk(4, dim=2)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g); h = torch.sigmoid(o) * torch.tanh(c); out.append(h)
    Hs = torch.stack(out, 0)
    (Hs * d["dH"].double()).sum().backward()
    return Hs.detach(), xw.grad


if __name__ == "__main__":
    for (T, K, N) in ((40, 2, 77), (33, 2, 32), (5, 1, 1)):
        d = make(T, K, N, seed=T)
        a, b = run(0, d, T, K, N), run(1, d, T, K, N)
        Hr, dzr = reference(d, T, K, N)
        for k in a:
            print(f"T={T} K={K} N={N} {k:6s} max|mma-fma| = {float((a[k] - b[k]).abs().max()):.3e}   scale {float(b[k].abs().max()):.3e}")
        for name, r in (("mma", a), ("fma", b)):
            print(f"   vs float64 autograd [{name}]: Hs {float((r['Hs'].double() - Hr).abs().max()):.3e}  dz {float((r['dz'].double() - dzr).abs().max()):.3e} (scale {float(dzr.abs().max()):.3e})")
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 750
    K = 2
    d = make(T, K, N)
    gates = torch.empty(T, K, N, 192, device=dev); Cs = torch.empty(T, K, N, 48, device=dev); Hs = torch.empty_like(Cs); HM = torch.empty_like(Cs)
    dz = torch.empty_like(gates); db = torch.empty(L.irrl_lstm_seq_ctas(N), K, 192, device=dev)
    gb_f = (2 * gates.numel() + 3 * Cs.numel()) * 4 / 1e9; gb_b = (2 * gates.numel() + 3 * Cs.numel()) * 4 / 1e9
    for path, name in ((0, "tensor-core"), (1, "fp32-fma"), (0, "tensor-core")):
        L.irrl_lstm_seq_set_path(path)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for rep in range(2):
            ev[0].record()
            _lib.check(L.irrl_lstm_seq_fwd(st(), T, K, N, p(d["xw"]), p(d["wh"]), p(d["c0"]), p(d["h0"]), p(d["keep"]), p(gates), p(Cs), p(Hs), p(d["b"]),
                                           None if os.environ.get("NOHM") else p(HM)))       # NOHM=1: as the learner calls it (no masked copy of h)
            ev[1].record()
            _lib.check(L.irrl_lstm_seq_bwd(st(), T, K, N, p(d["dH"]), p(d["wh"]), p(d["c0"]), p(d["keep"]), p(gates), p(Cs), p(dz), p(db)))
            ev[2].record(); torch.cuda.synchronize()
        f, b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        print(f"{name:12s} N={N} T={T}: fwd {f:7.2f} ms ({gb_f / f * 1e3:6.0f} GB/s algorithmic)   bwd {b:7.2f} ms ({gb_b / b * 1e3:6.0f} GB/s)")
