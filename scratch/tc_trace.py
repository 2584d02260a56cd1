import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from normflow__b200 import _C, _ops
dev = 'cuda'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K, L0, L1 = 10, 64, 64
P = 3 * K - 2
torch.manual_seed(0)
w = [torch.randn(8, 1, 3, 3, device=dev) / 3, torch.randn(8, 8, 3, 3, device=dev) / 72 ** 0.5, torch.randn(P, 8, 3, 3, device=dev) / 72 ** 0.5]
x = torch.randn(B, L0, L1, device=dev)
prm = _C.RqsParams(K, -5.0, 5.0, -5.0, 5.0, 1, 1)
trace = torch.zeros(8192, dtype=torch.int64, device=dev)
lib = _C.lib()
lib.nfk_debug_tc_trace.argtypes = [ctypes.c_void_p]
lib.nfk_debug_tc_trace.restype = None
with torch.no_grad():
    for _ in range(2):
        _ops.fused2d_step(x, w, [None] * 3, 1, prm, 0, 0)
    torch.cuda.synchronize()
    lib.nfk_debug_tc_trace(trace.data_ptr())
    _ops.fused2d_step(x, w, [None] * 3, 1, prm, 0, 0)
    torch.cuda.synchronize()
    lib.nfk_debug_tc_trace(None)
t = trace.cpu().numpy()
c = t[:4096]; m = t[4096:]
nc = int((c > 0).sum()); nm = int((m > 0).sum())
c = c[:nc].reshape(-1, 6); m = m[:nm].reshape(-1, 5)
t0 = c[0, 0]
print("compute stamps per unit: [start, E2 done, x stored+bar, P1 done, E3 done, end]   (cycles, relative)")
for i in range(3, 13):
    print(i, (c[i] - c[i, 0]).tolist(), " unit period", int(c[i + 1, 0] - c[i, 0]))
print("MMA stamps per unit: [wait H1, H1 ok, M2 issued, H2 ok, M3 issued]")
for i in range(3, 13):
    print(i, (m[i] - c[i, 0]).tolist())
d = np.diff(c[:, 0])
print("mean unit period", d[5:].mean(), " phases mean:", (c[5:, 1:] - c[5:, :-1]).mean(0).tolist())
print("MMA: wait for H1", (m[5:, 1] - m[5:, 0]).mean(), " M2 issue", (m[5:, 2] - m[5:, 1]).mean(), " wait H2", (m[5:, 3] - m[5:, 2]).mean(), " M3 issue", (m[5:, 4] - m[5:, 3]).mean())
