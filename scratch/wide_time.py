"""Wide conditioners ConvAct(1 -> H -> H -> 28) on the 64 x 64 lattice ("report a H = 64 variant as the real
dense GEMM case"): one RQ-spline coupling step through nfk_fusednd_step vs the layer-by-layer CUDA-core kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from normflow__b200 import _ops, _C
from normflow__b200.mask import EvenOddMask
shape, K, P = (64, 64), 10, 28
B = int(os.environ.get("BW", 4096))
prm = _C.RqsParams(K, -5.0, 5.0, -5.0, 5.0, 1, 1)
mask = EvenOddMask(shape=shape).to('cuda')._mask
for H in (16, 32, 64):
    g = torch.Generator('cpu').manual_seed(0)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device='cpu') * sc).cuda()
    w = [rnd(H, 1, 3, 3, sc=0.3), rnd(H, H, 3, 3, sc=0.5 / (9 * H) ** 0.5), rnd(P, H, 3, 3, sc=0.5 / (9 * H) ** 0.5)]
    x = rnd(B, *shape, sc=1.2)
    def tc():
        return _ops.fusednd_step(x, w, [None] * 3, 1, prm, 0, 0, 0, False)
    def layerwise():
        out = _ops.conv_stack(x.unsqueeze(1), w, [None] * 3, ('tanh', 'tanh', None), 3, in_mask=mask, in_keep=0)
        return _ops.rqs_apply(x, out, mask, 0, prm, 0, _C.FROZEN_COPY, False)
    res = {}
    for name, fn in (("tensor-core", tc), ("layer-by-layer", layerwise)):
        with torch.no_grad():
            for _ in range(2): y = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): y = fn()
            e1.record(); torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / 3, y)
    V = 4096
    flop = 2 * 9 * (H + H * H + H * P) * V * B                     # fp32-equivalent flops of the conditioner
    t_tc, t_lw = res["tensor-core"][0], res["layer-by-layer"][0]
    dy = float((res["tensor-core"][1][0] - res["layer-by-layer"][1][0]).abs().max())
    print(f"H={H} B={B}: tensor-core {t_tc:.3f} ms/step ({B / t_tc * 1e3 / 4:.0f} samples/s for 4 steps, "
          f"{flop / t_tc * 1e-9:.1f} TFLOP/s fp32-equivalent) | layer-by-layer {t_lw:.3f} ms/step | x{t_lw / t_tc:.1f} | max |dy| {dy:.1e}", flush=True)
