import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['NFK_WGRAD_TC'] = '1'; os.environ['NFK_WGRAD_TRACE'] = '1'
import numpy as np, torch
import normflow__b200 as nf
from normflow__b200 import _ops, _C
B, L0, L1, Co, parity = 4096, 64, 64, 28, 1
x = torch.tanh(torch.randn(B, 8, L0, L1, device='cuda'))
g = (torch.randn(B, Co, L0, L1, device='cuda') * 1e-4).contiguous()
for _ in range(2):
    _ops._conv_weight_grad(x, None, 0, g, (Co, 8, 3, 3), False, (L0, L1), 3, parity)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 2048)()
_C.lib().nfk_debug_wgrad_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
_C.lib().nfk_debug_wgrad_trace(buf, 2048)
t = np.array(buf[:], dtype=np.int64)
p = t[:1000]; n = (p > 0).sum() // 7
p = p[:n * 7].reshape(n, 7)
names = ['wait_group', 'bar_producers', 'issue', 'mbar_empty', 'tasks', 'fence+arrive', 'loop/drain']
d = np.diff(p, axis=1)
print("producer thread 160, cycles per phase (median over tiles 20..%d):" % n)
for k in range(6):
    print(f"  {names[k]:>14s}: {np.median(d[20:, k]):8.0f}")
print(f"  {names[6]:>14s}: {np.median(p[21:, 0] - p[20:-1, 6]):8.0f}")
print("  period:", np.median(np.diff(p[20:, 0])))
m = t[1024:1024 + 1000].reshape(250, 4)
print("MMA lead: wait full %.0f, issue %.0f, period %.0f" % (np.median(m[20:, 1] - m[20:, 0]), np.median(m[20:, 2] - m[20:, 1]),
                                                         np.median(np.diff(m[20:, 0]))))
