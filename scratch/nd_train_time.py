"""Training-step time at config-4 / config-5 geometry with per-kernel-family timers (NFK_DGRAD_TC=0/1 A/B)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_configs as BC
import normflow__b200
from normflow__b200 import _C
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
timer = _C.KernelTimer(); _C.kernel_timer = timer
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    fit.step()
e1.record(); torch.cuda.synchronize()
_C.kernel_timer = None
ms = e0.elapsed_time(e1) / n
print(f"config {cfgno} train step B={B} (NFK_DGRAD_TC={os.environ.get('NFK_DGRAD_TC', '1')}): {ms:.1f} ms -> {B / ms * 1e3:.0f} samples/s")
for k, v in sorted(timer.summary().items(), key=lambda kv: -kv[1]['total_ms'])[:12]:
    print(f"   {k:32s} {v['launches']:4d} launches  {v['total_ms'] / n:8.2f} ms/step")
