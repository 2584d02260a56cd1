"""Training-step and MCMC throughput of the bench workload (config 3) -- secondary numbers for DESIGN.md."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from normflow__b200 import _C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = bench.build_model(torch)
fit = model.fit
# one epoch through the public API to build the optimizer, then time fit.step()
model.fit(n_epochs=1, batch_size=B, checkpoint_dict=dict(print_stride=1000, print_batch_size=64, display=False))
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
print(f"train step B={B}: {ms:.1f} ms -> {B / ms * 1e3:.0f} samples/s; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
for k, v in sorted(timer.summary().items(), key=lambda kv: -kv[1]['total_ms']):
    print(f"   {k:32s} {v['launches']:4d} launches  {v['total_ms'] / n:8.2f} ms/step")
np.random.seed(0)
model.mcmc.sample(B)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    model.mcmc.sample(B)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"mcmc.sample B={B}: {dt * 1e3:.1f} ms -> {B / dt:.0f} samples/s")
