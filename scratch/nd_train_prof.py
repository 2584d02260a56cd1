"""One training step at config-4 / config-5 geometry (for an ncu launch list: where does the N-D backward go?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_configs as BC
import normflow__b200
cfgno, B = int(sys.argv[1]), int(sys.argv[2])
model = BC.build_model(normflow__b200, BC.CONFIGS[cfgno])
model.device_handler.to('cuda')
fit = model.fit
fit.loss_fn = fit.calc_kl_mean
fit.optimizer = torch.optim.AdamW(model.net_.parameters(), lr=1e-3, fused=True)
fit.train_batch_size = B
for _ in range(2):
    fit.step()
torch.cuda.synchronize()
