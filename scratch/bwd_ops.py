"""The tensor-core backward ops in isolation (for ncu launch lists): data and weight gradient of the 8 -> Co layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import normflow__b200
from normflow__b200 import _ops, _C
shape = tuple(int(v) for v in sys.argv[1].split('x'))
B, Co = int(sys.argv[2]), int(sys.argv[3])
os.environ['NFK_WGRAD_ND_TC'] = '1'
os.environ['NFK_WGRAD_TC'] = '0'
os.environ['NFK_DGRAD_TC'] = '1'
dev = 'cuda'
h = torch.tanh(torch.randn(B, 8, *shape, device=dev))
gpre = torch.randn(B, Co, *shape, device=dev) * 1e-3
sparse = len(sys.argv) > 4 and sys.argv[4] == 'sparse'
if sparse:
    from normflow__b200.mask import EvenOddMask
    gpre = gpre * EvenOddMask(shape=shape)._mask.to(dev)          # non-zero where the coordinate sum is even
w = torch.randn(Co, 8, *(3,) * len(shape), device=dev) * 0.1
for _ in range(3):
    gw, gb = _ops._conv_weight_grad(h, None, 0, gpre, tuple(w.shape), True, shape, 3)
    gin = _ops._conv_dgrad_tc(gpre, w, h, _C.ACT['tanh'], shape, 3, g_parity=0 if sparse else None)
torch.cuda.synchronize()
