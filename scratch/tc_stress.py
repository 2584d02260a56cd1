import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from normflow__b200 import _C, _ops
import test_gpu_parity as T
DEV = 'cuda'
K, kind, shape, B, P = 10, 1, (32, 32), 24, 28
prm = _C.RqsParams(K, -5.0, 5.0, -5.0, 5.0, 1, 1)
for scale in (1.0, 2.0, 4.0, 8.0):
    g = torch.Generator('cpu').manual_seed(11)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device='cpu') * sc).to(DEV)
    w = [rnd(8, 1, 3, 3, sc=0.3 * scale), rnd(8, 8, 3, 3, sc=scale * 0.5 / 72 ** 0.5), rnd(P, 8, 3, 3, sc=scale * 0.5 / 72 ** 0.5)]
    b = [None] * 3
    x = rnd(B, *shape, sc=1.3)
    yo, lo = T._oracle_single_step(x, w, b, kind, 0, K, False, 0)
    for tc in ('1', '0'):
        os.environ['NFK_FUSED_TC'] = tc
        with torch.no_grad():
            y, lj = _ops.fused2d_step(x, w, b, kind, prm, 0, 0, 0, False)
        dy = np.abs(y.double().cpu().numpy() - yo) / np.maximum(np.abs(yo), 1) / 1e-5
        dl = np.abs(lj.double().cpu().numpy() - lo) / np.maximum(np.abs(lo), 1) / 1e-5
        print(f"weight scale x{scale}: tc={tc}: y excess max {dy.max():.2f} p99.9 {np.quantile(dy, 0.999):.2f}; logJ excess max {dl.max():.2f}  (|logJ| median {np.median(np.abs(lo)):.1f})")
