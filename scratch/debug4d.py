import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import _config_model, _oracle_flow, ACTION, O
for shape, blocks in [((16,16,16,16), [('affine', 2)]), ((16,16,16,16), [('rqs', 2)]), ((8,8,8,8), [('rqs', 2)]), ((32,32,32), [('rqs',2)])]:
    for scale in (1.0, 1.7):
        model = _config_model(shape, blocks, seed=1)
        with torch.no_grad():
            for p in model.net_.parameters(): p.mul_(scale)
            x = model.prior.sample(2)
            y, lj = model.net_(x)
            xb, lb = model.net_.backward(y, log0=lj)
            print(shape, blocks, scale, "fwd finite", torch.isfinite(y).all().item(), "roundtrip", (xb-x).abs().max().item(), "res log", lb.abs().max().item(), "|lj|", lj.abs().max().item(), flush=True)
            if scale == 1.7 and shape[0] <= 16:
                yr, lr = _oracle_flow(model, x.cpu().numpy())
                print("   vs oracle: y", np.abs(y.double().cpu().numpy()-yr).max(), "logJ", np.abs(lj.double().cpu().numpy()-lr).max(), flush=True)
