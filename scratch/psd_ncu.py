"""One PSD-block evaluation at config-3 size for an ncu capture of nfk_psd_scale."""
import numpy as np, torch
import normflow__b200 as nf
from normflow__b200.nn import FFTNet_, MeanFieldNet_, PSDBlock_
lat, B = (64, 64), 16384
blk = PSDBlock_(mfnet_=MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True),
                fftnet_=FFTNet_.build(lat, knots_len=10, ignore_zeromode=True))
x = torch.randn(B, *lat, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y, l = blk(x)
torch.cuda.synchronize()
print(float(y.abs().mean()), float(l.mean()))
