"""PSD block timings on one B200: the spectral-scale kernel against its HBM model, cuFFT's share,
the whole block, and the reference's op sequence run with torch CUDA ops for comparison."""
import json
import sys
import numpy as np
import torch
import normflow__b200 as nf
from normflow__b200 import _ops
from normflow__b200.nn import FFTNet_, MeanFieldNet_, PSDBlock_

dev = "cuda"
out = {}


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for lat, B in [((64, 64), 16384), ((32, 32, 32), 512), ((16, 16, 16, 16), 256), ((16, 16), 65536)]:
    torch.manual_seed(0)
    blk = PSDBlock_(mfnet_=MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True),
                    fftnet_=FFTNet_.build(lat, knots_len=10, ignore_zeromode=True))
    ff = blk.fftnet_
    x = torch.randn(B, *lat, device=dev)
    V = int(np.prod(lat))
    res = {}
    with torch.no_grad():
        spec = ff.spectrum(x)
        w, _ = ff.weights()
        zm = torch.randn(B, device=dev)
        Kc = w.numel()
        t = timeit(lambda: _ops.psd_scale(spec, w, zm, float(V)))
        res["psd_scale_ms"] = t
        res["psd_scale_GBps"] = B * Kc * 16 / t / 1e6
        out_buf = torch.empty_like(spec)
        res["rfftn_ms"] = timeit(lambda: ff.spectrum(x))
        res["irfftn_ms"] = timeit(lambda: ff.field(spec))
        res["block_ms"] = timeit(lambda: blk(x))
        res["block_samples_per_s"] = B / res["block_ms"] * 1e3
        res["block_GBps_8B_per_site_model"] = B * V * 8 / res["block_ms"] / 1e6

        def literal():
            dim = list(range(1, x.dim()))
            xm = torch.mean(x, dim=dim).reshape(-1, *[1 for _ in dim])
            ymf, l1 = blk.mfnet_.forward(xm, rvol=V ** 0.5)
            ww = 1 / ff.ipsd ** 0.5
            y = torch.fft.irfftn(torch.fft.rfftn(x - xm, dim=ff.rfft_dim) * ww, dim=ff.rfft_dim)
            return ymf + y, l1 + ff.log_jacobian(ww)
        res["literal_torch_ops_ms"] = timeit(literal)
    xg = x.clone().requires_grad_(True)

    def train():
        y, l = blk(xg)
        (y.square().sum() + l.sum()).backward()
    res["fwd_bwd_ms"] = timeit(train, n=10, warm=3)
    out[f"{'x'.join(map(str, lat))}_B{B}"] = res
    print(lat, B, json.dumps(res), flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/psd_time.json", "w"), indent=1)
