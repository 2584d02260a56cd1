import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from normflow__b200 import _C, _ops
from oracle import nf_oracle as O
import test_gpu_parity as T
DEV = 'cuda'
g = torch.Generator('cpu').manual_seed(11)
K, kind, shape, B = 10, 1, (16, 16), 700
P = 28
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device='cpu') * scale).to(DEV)
w = [rnd(8, 1, 3, 3, scale=0.3), rnd(8, 8, 3, 3, scale=0.5 / 72 ** 0.5), rnd(P, 8, 3, 3, scale=0.5 / 72 ** 0.5)]
b = [None] * 3
x = rnd(B, *shape, scale=1.3)
prm = _C.RqsParams(K, -5.0, 5.0, -5.0, 5.0, 1, 1)
for parity in (0, 1):
    yo, lo = T._oracle_single_step(x, w, b, kind, parity, K, False, 0)
    for tc in ('1', '0'):
        os.environ['NFK_FUSED_TC'] = tc
        with torch.no_grad():
            y, lj = _ops.fused2d_step(x, w, b, kind, prm, 0, parity, 0, False)
        dy = np.abs(y.double().cpu().numpy() - yo) / np.maximum(np.abs(yo), 1)
        dl = np.abs(lj.double().cpu().numpy() - lo)
        print(f"parity {parity} tc={tc}: y max excess {dy.max()/1e-5:.2f}  logJ abs err max {dl.max():.2e} rms {np.sqrt((dl**2).mean()):.2e}  |logJ| median {np.median(np.abs(lo)):.2f}  excess {np.max(dl/(1e-5*np.maximum(np.abs(lo),1))):.2f}")
