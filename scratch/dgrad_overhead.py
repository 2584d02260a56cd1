"""Where do the un-timed milliseconds of a training step with the tensor-core data gradient go? (host wall time per call)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from normflow__b200 import _ops, _C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = bench.build_model(torch)
fit = model.fit
model.fit(n_epochs=1, batch_size=B, checkpoint_dict=dict(print_stride=1000, print_batch_size=64, display=False))
torch.cuda.synchronize()
acc = {}
orig_empty = torch.empty
def timed_empty(*a, **k):
    t = time.perf_counter(); r = orig_empty(*a, **k); dt = time.perf_counter() - t
    n = r.numel() * r.element_size()
    key = 'empty>=256MB' if n >= 2**28 else 'empty<256MB'
    acc[key] = acc.get(key, 0) + dt; acc[key + '_n'] = acc.get(key + '_n', 0) + 1
    return r
torch.empty = timed_empty
orig = _ops._conv_dgrad_tc
def wrapped(*a, **k):
    t = time.perf_counter(); r = orig(*a, **k); acc['dgrad_host'] = acc.get('dgrad_host', 0) + time.perf_counter() - t
    return r
_ops._conv_dgrad_tc = wrapped
for trial in range(2):
    acc.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        fit.step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"3 steps: host enqueue {1e3 * (t1 - t0):.1f} ms, until idle {1e3 * (t2 - t0):.1f} ms;", {k: round(v * 1e3, 2) if not k.endswith('_n') else v for k, v in acc.items()})
print(torch.cuda.memory_stats()['num_alloc_retries'], torch.cuda.memory_stats()['segment.all.allocated'], torch.cuda.memory_stats()['reserved_bytes.all.peak'] / 2**30)
