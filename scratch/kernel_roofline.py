"""Per-kernel roofline at BASELINE config-3 sizes: algorithmic bytes / CUDA-event time
against the measured HBM copy peak.  Writes profiles/r02_kernel_roofline.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from normflow__b200 import _C, _ops
from normflow__b200.mask import EvenOddMask
dev = 'cuda'
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
B, L = 16384, (64, 64)
V = L[0] * L[1]
K, P = 10, 28
torch.manual_seed(0)
x = torch.randn(B, *L, device=dev)
out28 = torch.randn(B // 4, P, *L, device=dev) * 0.5            # quarter batch: 1.9 GB tensor
x4 = x[:B // 4].contiguous()
out2 = torch.randn(B, 2, *L, device=dev) * 0.5
mask = EvenOddMask(shape=L).to(dev)._mask
prm = _ops.rqs_params(K, (-5, 5), (-5, 5), dict(left='linear', right='linear'))
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()                                  # 256 MB > L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

rows = []
def add(name, fn, nbytes, note=""):
    ms = timeit(fn)
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows.append(dict(kernel=name, ms=round(ms, 4), algorithmic_bytes=int(nbytes), achieved_gbs=round(gbs, 1),
                     frac_of_measured_hbm_peak=round(gbs / peak, 3), note=note))
    print(f"{name:34s} {ms:9.4f} ms  {gbs:8.1f} GB/s  {gbs / peak:6.1%}  {note}")

with torch.no_grad():
    add("prior_normal_sample (+log r)", lambda: _ops.prior_sample(B, L, None, None, 1, 0, dev), 4 * B * V, "write x")
    add("prior_normal_logprob", lambda: _ops.prior_logprob(x, None, None), 4 * B * V, "read x")
    add("phi4_action_fwd", lambda: _ops.phi4_action(x, 0.67, 0.0, 0.5), 4 * B * V, "read phi")
    add("affine_fwd", lambda: _ops.affine_apply(x, out2, mask, 0), 16 * B * V, "x, t, s, y")
    add("rqs_fwd (K=10)", lambda: _ops.rqs_apply(x4, out28, mask, 0, prm), (8 + 4 * P) * (B // 4) * V, "x, 28 params, y; B/4")
    add("rqs_inv (K=10)", lambda: _ops.rqs_apply(x4.clamp(-4.9, 4.9), out28, mask, 0, prm, inverse=True), (8 + 4 * P) * (B // 4) * V, "B/4 (incl. clamp kernel)")
    add("mask_select", lambda: _ops.mask_select(x, mask, 1), 8 * B * V, "x, y")
    idx = torch.randint(0, B, (B,), device=dev)
    add("gather_rows", lambda: _ops.gather_rows(x, idx), 8 * B * V, "read + write rows")
    xs = torch.randn(B, *L, device=dev) * 0.3
    from normflow__b200.nn import DistConvertor_
    dc = DistConvertor_(10, symmetric=True).to(dev)
    add("distconvertor chain (spline1d)", lambda: dc(xs), 8 * B * V, "x, y")
    logq, logp = torch.randn(B, device=dev), torch.randn(B, device=dev)
    lu = torch.log(torch.rand(B, device=dev, dtype=torch.float64))
    ref = torch.zeros(2, device=dev, dtype=torch.float64)
    add("metropolis_scan (B=16384)", lambda: _ops.metropolis_scan(logq, logp, lu, ref), 24 * B, "latency-bound sequential scan")
# backward kernels (need autograd objects): time the raw C calls through the autograd functions
xg = x4.clone().requires_grad_(True)
og = out28.clone().requires_grad_(True)
y, lj = _ops.rqs_apply(xg, og, mask, 0, prm)
gy, gl = torch.randn_like(y), torch.randn_like(lj)
def rqs_bwd():
    torch.autograd.grad([y, lj], [xg, og], [gy, gl], retain_graph=True)
add("rqs_bwd (K=10)", rqs_bwd, (12 + 8 * P) * (B // 4) * V, "x, params, gy -> gx, gparams; B/4")
xa = x.clone().requires_grad_(True)
S = _ops.phi4_action(xa, 0.67, 0.0, 0.5)
gS = torch.randn_like(S)
add("phi4_action_bwd", lambda: torch.autograd.grad(S, xa, gS, retain_graph=True), 12 * B * V, "phi, gphi (+ halo re-reads in L2)")
json.dump(dict(round=1, hbm_peak_gbs=peak, batch=B, lattice=list(L), rows=rows),
          open(os.path.join(ROOT, 'gpurun_out', 'kernel_roofline.json'), 'w'), indent=1)
