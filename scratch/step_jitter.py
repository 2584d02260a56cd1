import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
model = bench.build_model(torch)
B = 16384
for _ in range(3): model.posterior.sample__(B)
torch.cuda.synchronize()
for trial in range(3):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    t0 = time.perf_counter(); host = []
    evs[0].record()
    for i in range(10):
        y, logq, logp = model.posterior.sample__(B)
        evs[i + 1].record(); host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    print("trial", trial, "gpu ms/step:", [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(10)])
    print("        host enqueue done at ms:", [round(h * 1e3, 2) for h in host], " mem GiB", round(torch.cuda.memory_reserved() / 2**30, 2))
