"""One RQ-spline and one affine coupling step of the N-D tensor-core path at config-4 / config-5 geometry
(for an ncu launch list: per-kernel durations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from normflow__b200 import _ops, _C
def run(shape, B, kind, K=10):
    D = len(shape)
    P = 2 if kind == 0 else 3 * K - 2
    g = torch.Generator('cpu').manual_seed(0)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device='cpu') * sc).cuda()
    fan = 8 * 3 ** D
    w = [rnd(8, 1, *(3,) * D, sc=0.3), rnd(8, 8, *(3,) * D, sc=0.5 / fan ** 0.5), rnd(P, 8, *(3,) * D, sc=0.5 / fan ** 0.5)]
    x = rnd(B, *shape, sc=1.2)
    prm = _C.RqsParams(K, -5.0, 5.0, -5.0, 5.0, 1, 1) if kind == 1 else None
    for _ in range(2):
        with torch.no_grad():
            y, l = _ops.fusednd_step(x, w, [None] * 3, kind, prm, 0, 0, 0, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        y, l = _ops.fusednd_step(x, w, [None] * 3, kind, prm, 0, 0, 0, False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    V = x[0].numel()
    print(f"{shape} B={B} kind={kind}: {ms:.3f} ms/step -> {ms * 1e-3 * 1.965e9 * 148 / (B * V):.1f} SM-cycles per site", flush=True)
b4, b5 = int(os.environ.get("B4", 512)), int(os.environ.get("B5", 64))
run((32, 32, 32), b4, 1); run((32, 32, 32), b4, 0)
run((16, 16, 16, 16), b5, 1); run((16, 16, 16, 16), b5, 0)
