"""Times the previous build's N-D step (scratch/libnfk_old.so, old workspace signature) for an A/B in one session."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from normflow__b200 import _C
from normflow__b200._C import dev, stream, Lattice, RqsParams, c_f, c_i, c_l
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnfk_old.so"))
lib.nfk_fusednd_workspace.argtypes = [Lattice, c_i, c_i, c_l]; lib.nfk_fusednd_workspace.restype = c_l
lib.nfk_fusednd_step.argtypes = [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, RqsParams, Lattice, c_i, c_i, c_i, c_f, c_f, c_f, c_l, c_f, c_l, c_f]
def run(shape, B, kind, K=10):
    D = len(shape); P = 2 if kind == 0 else 3 * K - 2
    g = torch.Generator('cpu').manual_seed(0)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device='cpu') * sc).cuda()
    fan = 8 * 3 ** D
    w = [rnd(8, 1, *(3,) * D, sc=0.3), rnd(8, 8, *(3,) * D, sc=0.5 / fan ** 0.5), rnd(P, 8, *(3,) * D, sc=0.5 / fan ** 0.5)]
    x = rnd(B, *shape, sc=1.2); y = torch.empty_like(x); lo = torch.empty(B, device='cuda')
    prm = RqsParams(K, -5.0, 5.0, -5.0, 5.0, 1, 1) if kind == 1 else RqsParams(2, 0.0, 1.0, 0.0, 1.0, 0, 0)
    lat = _C.lattice(shape)
    need = lib.nfk_fusednd_workspace(lat, kind, prm.n_knots, B)
    ws = torch.empty(need, dtype=torch.uint8, device='cuda')
    def call():
        rc = lib.nfk_fusednd_step(dev(x), dev(w[0]), None, dev(w[1]), None, dev(w[2]), None, 8, kind, prm, lat, 0, 0, 0, None, dev(y), dev(lo), B, dev(ws, torch.uint8), need, stream())
        assert rc == 0, rc
    for _ in range(2): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"OLD {shape} B={B} kind={kind}: {ms:.3f} ms/step -> {ms * 1e-3 * 1.965e9 * 148 / (B * x[0].numel()):.1f} SM-cycles per site", flush=True)
run((32, 32, 32), 512, 1); run((32, 32, 32), 512, 0); run((16,) * 4, 64, 1); run((16,) * 4, 64, 0)
