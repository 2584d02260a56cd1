"""Config-5 stack: forward and inverse, block by block, N-D tensor-core path vs the layer-by-layer float32 kernels
(NFK_FUSED_ND=0), and both against the float64 oracle in the forward direction."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import test_gpu_parity as T
shape, blocks = (16,) * 4, [('affine', 4), ('rqs', 4)] * 2
model = T._config_model(shape, blocks)
x = torch.randn(2, *shape, generator=torch.Generator('cpu').manual_seed(1234), dtype=torch.float32, device='cpu').cuda()
def run(flag, inp, inverse):
    os.environ['NFK_FUSED_ND'] = flag
    outs = []
    v, l = inp, torch.zeros(inp.shape[0], device='cuda')
    with torch.no_grad():
        seq = list(model.net_)
        for cpl in (reversed(seq) if inverse else seq):
            v, l = (cpl.backward(v, l) if inverse else cpl(v, l))
            outs.append((v.clone(), l.clone()))
    return outs
f1, f0 = run('1', x, False), run('0', x, False)
for i, (a, b) in enumerate(zip(f1, f0)):
    print(f"fwd block {i}: tc vs layerwise max |dy| {float((a[0]-b[0]).abs().max()):.2e}  |dlogJ| {float((a[1]-b[1]).abs().max()):.2e}")
y = f0[-1][0]
i1, i0 = run('1', y, True), run('0', y, True)
for i, (a, b) in enumerate(zip(i1, i0)):
    d = (a[0]-b[0]).abs()
    print(f"inv block {i}: tc vs layerwise max |dx| {float(d.max()):.2e} (count > 1e-3: {int((d > 1e-3).sum())})  |dlog| {float((a[1]-b[1]).abs().max()):.2e}")
print("round trip layerwise:", float((i0[-1][0] - x).abs().max()), " tc:", float((i1[-1][0] - x).abs().max()))
# same-path round trips
y1 = f1[-1][0]
j1 = run('1', y1, True)
print("tc forward -> tc inverse:", float((j1[-1][0] - x).abs().max()))
d = (j1[-1][0] - x).abs().flatten()
top = torch.topk(d, 5)
print("worst sites", top.values.tolist(), top.indices.tolist())
