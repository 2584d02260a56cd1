"""Times the UNMODIFIED reference (oracle/_ref/normflow_ref, staged by oracle/stage_ref.py) on the
host cores  --  TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

    CUDA_VISIBLE_DEVICES="" python -m oracle.ref_runner --config 3 --what sample --batch 256 --steps 4 --warmup 1

prints one JSON object {"config", "what", "samples_per_s", "seconds", "batch", "steps", "threads",
"dtype": "f64", "kind": "reference"}.  `what`: sample = `model.posterior.sample__(B)` (the
BASELINE metric: flow forward + log|det J| + action, src/_normflowcore.py:109-119), train =
`model.fit.step()` (:275-294, AdamW), mcmc = `model.mcmc.sample(B)` (src/mcmc/mcmc.py:40-87).
The reference runs in its own default float64 on `threads` ATen threads (all host cores).

bench.py runs this module in a child process with the GPUs hidden, because the reference makes
CUDA its default device at import when one is visible (src/device/__init__.py:7-13).
"""

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def time_reference(config, what, batch, steps, warmup):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""                 # must precede the first CUDA query
    from oracle import stage_ref
    ref = stage_ref.import_reference()
    import bench_configs as C
    torch = ref.torch
    if torch.cuda.is_available():
        raise RuntimeError("ref_runner: a GPU is visible; run with CUDA_VISIBLE_DEVICES=\"\"")
    threads = host_threads()
    torch.set_num_threads(threads)
    cfg = C.CONFIGS[config]
    model = C.build_model(ref, cfg)
    import numpy as np
    np.random.seed(7)
    if what == "train":
        fit = model.fit
        fit.optimizer = torch.optim.AdamW(model.net_.parameters(), lr=1e-3, weight_decay=0.01)
        fit.loss_fn = fit.calc_kl_mean
        fit.scheduler = None
        fit.train_batch_size = batch
        one = fit.step
    elif what == "mcmc":
        one = lambda: model.mcmc.sample(batch)
    else:
        one = lambda: model.posterior.sample__(batch)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return dict(config=config, what=what, samples_per_s=batch * steps / dt, seconds=dt, batch=batch,
                steps=steps, threads=threads, dtype="f64", kind="reference")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", default=None,
                    help='JSON list of [config, what, batch, steps, warmup]: one JSON line per job, in order')
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--what", default="sample", choices=["sample", "train", "mcmc"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    if a.jobs:
        for config, what, batch, steps, warmup in json.loads(a.jobs):
            try:
                out = time_reference(int(config), what, int(batch), int(steps), int(warmup))
            except Exception as err:                      # one failed job must not lose the others
                out = dict(config=config, what=what, error=f"{type(err).__name__}: {err}")
            print(json.dumps(out), flush=True)
        return
    print(json.dumps(time_reference(a.config, a.what, a.batch, a.steps, a.warmup)), flush=True)


if __name__ == "__main__":
    main()
