"""Where does the float32 error of the full config-5 stack (16 coupling steps, Conv4d) come from?
Per coupling block: cumulative error of the GPU flow against the float64 oracle, and the error of that block ALONE
when the oracle is fed the GPU's own input (isolates per-block rounding from amplification of earlier error)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import test_gpu_parity as T
from oracle import nf_oracle as O
shape, blocks = (16,) * 4, [('affine', 4), ('rqs', 4)] * 2
if len(sys.argv) > 1 and sys.argv[1] == '3d':
    shape, blocks = (32,) * 3, [('affine', 4), ('rqs', 4)]
model = T._config_model(shape, blocks)
x = torch.randn(1, *shape, generator=torch.Generator('cpu').manual_seed(1234), dtype=torch.float32, device='cpu')
ex = lambda got, ref: float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1)) / 1e-5)
with torch.no_grad():
    stack = model.net_.hack(x.cuda(), log0=0)
y_or = x.numpy().astype(np.float64)
for bi, cpl in enumerate(model.net_):
    sub = type('M', (), {})()
    sub.net_ = [cpl]
    y_in_gpu = stack[bi][0].double().cpu().numpy()
    y_out_gpu = stack[bi + 1][0].double().cpu().numpy()
    y_or, _ = T._oracle_flow(sub, y_or)                 # cumulative oracle
    y_iso, _ = T._oracle_flow(sub, y_in_gpu)            # this block alone, fed the GPU's input
    print(f"block {bi} ({blocks[bi][0]} x{blocks[bi][1]}): cumulative excess {ex(y_out_gpu, y_or):6.2f}   "
          f"block alone {ex(y_out_gpu, y_iso):6.2f}   |y| max {np.abs(y_or).max():.2f}", flush=True)
