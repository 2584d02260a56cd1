import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['NFK_WGRAD_TC'] = '0'
import torch
import normflow__b200 as nf
from normflow__b200 import _ops
B, L0, L1, Co = 4096, 64, 64, 8
x = torch.tanh(torch.randn(B, 8, L0, L1, device='cuda'))
g = (torch.randn(B, Co, L0, L1, device='cuda') * 1e-4).contiguous()
for _ in range(2):
    _ops._conv_weight_grad(x, None, 0, g, (Co, 8, 3, 3), False, (L0, L1), 3, None)
torch.cuda.synchronize()
