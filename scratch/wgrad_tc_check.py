"""Tensor-core weight gradient vs the CUDA-core kernels and a float64 torch reference; timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import normflow__b200 as nf
from normflow__b200 import _ops, _C

dev = "cuda"
torch.manual_seed(0)


def ref64(x, g):
    """gw[co,ci,kh,kw] = sum g[b,co,r,c] x[b,ci,r+kh-1,c+kw-1]"""
    x, g = x.double(), g.double()
    out = torch.zeros(g.shape[1], x.shape[1], 3, 3, dtype=torch.float64, device=x.device)
    for kh in range(3):
        for kw in range(3):
            xs = torch.roll(x, shifts=(1 - kh, 1 - kw), dims=(2, 3))
            out[:, :, kh, kw] = torch.einsum('bois,bcis->oc', g.unsqueeze(2).flatten(3), xs.unsqueeze(1).flatten(3)) \
                if False else torch.einsum('bors,bcrs->oc', g, xs)
    return out, g.sum(dim=(0, 2, 3))


def run(B, Co, L0, L1, parity, scale_g=1e-4, bias=True, check=True):
    x = torch.tanh(torch.randn(B, 8, L0, L1, device=dev))
    g = torch.randn(B, Co, L0, L1, device=dev) * scale_g * torch.exp(2 * torch.randn(B, Co, 1, 1, device=dev))
    if parity is not None:
        rr = torch.arange(L0, device=dev).view(-1, 1) + torch.arange(L1, device=dev).view(1, -1)
        g = g * ((rr % 2) == parity).float()
    g = g.contiguous()
    res = {}
    for mode in ("1", "0"):
        os.environ['NFK_WGRAD_TC'] = mode
        gw, gb = _ops._conv_weight_grad(x, None, 0, g, (Co, 8, 3, 3), bias, (L0, L1), 3, parity)
        torch.cuda.synchronize()
        res[mode] = (gw, gb)
    out = f"B={B} Co={Co} {L0}x{L1} parity={parity}:"
    if check:
        rw, rb = ref64(x, g)
        sw, sb = rw.abs().max().item(), rb.abs().max().item()
        for mode, name in (("1", "tc"), ("0", "cuda-core")):
            gw, gb = res[mode]
            ew = (gw.double() - rw).abs().max().item() / sw
            eb = (gb.double() - rb).abs().max().item() / max(sb, 1e-30) if bias else 0
            out += f"  {name}: rel err gw {ew:.2e} gb {eb:.2e}"
    print(out, flush=True)


def timeit(B, Co, L0, L1, parity):
    x = torch.tanh(torch.randn(B, 8, L0, L1, device=dev))
    g = (torch.randn(B, Co, L0, L1, device=dev) * 1e-4).contiguous()
    for mode in ("1", "0"):
        os.environ['NFK_WGRAD_TC'] = mode
        for _ in range(2):
            _ops._conv_weight_grad(x, None, 0, g, (Co, 8, 3, 3), False, (L0, L1), 3, parity)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            _ops._conv_weight_grad(x, None, 0, g, (Co, 8, 3, 3), False, (L0, L1), 3, parity)
        b.record(); torch.cuda.synchronize()
        print(f"   time B={B} Co={Co} {L0}x{L1} parity={parity} tc={mode}: {a.elapsed_time(b) / 5:.3f} ms", flush=True)


run(2, 28, 8, 8, 0)
run(3, 28, 8, 8, None)
run(5, 8, 16, 16, 1)
run(4, 2, 6, 12, 0)
run(7, 28, 64, 64, 1)
run(7, 8, 64, 64, None)
run(300, 28, 64, 64, 0)
run(2000, 28, 64, 64, 0)
run(2000, 8, 64, 64, None)
run(2, 28, 128, 128, 1)
run(2, 8, 10, 128, None)
run(3, 32, 2, 8, None, bias=False)
run(3, 28, 16, 16, 0)
run(2, 5, 7, 24, None)
timeit(4096, 28, 64, 64, 1)
timeit(4096, 8, 64, 64, None)
timeit(4096, 2, 16, 16, 0)
