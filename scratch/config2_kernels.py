"""Config 2 (16x16, affine x4, B=1024) training step: per-kernel GPU time (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_configs as BC
import normflow__b200
torch.manual_seed(0); np.random.seed(0)
model = BC.build_model(normflow__b200, BC.CONFIGS[2])
model.device_handler.to('cuda')
model.fit(n_epochs=3, batch_size=1024, checkpoint_dict=dict(print_stride=1000, display=False))
torch.cuda.synchronize()
model.fit.step()
torch.cuda.synchronize()
