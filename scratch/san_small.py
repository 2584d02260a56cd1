import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from normflow__b200 import _C, _ops
dev = 'cuda'
torch.manual_seed(0)
for shape, K, B in [((16, 16), 10, 5), ((6, 10), 5, 3), ((64, 64), 10, 2)]:
    P = 3 * K - 2
    w = [torch.randn(8, 1, 3, 3, device=dev) / 3, torch.randn(8, 8, 3, 3, device=dev) / 8, torch.randn(P, 8, 3, 3, device=dev) / 8]
    x = torch.randn(B, *shape, device=dev)
    prm = _C.RqsParams(K, -5.0, 5.0, -5.0, 5.0, 1, 1)
    with torch.no_grad():
        for parity in (0, 1):
            y, lj = _ops.fused2d_step(x, w, [None] * 3, 1, prm, 0, parity)
    torch.cuda.synchronize()
    print("ok", shape, float(y.abs().max()), float(lj.abs().max()))
print("san_small done")
