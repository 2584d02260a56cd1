#!/usr/bin/env python
"""bench.py -- samples/sec of "flow forward + log|det J| + action" on the 64x64 phi^4
workload of BASELINE.json (configs[2]: RQ-spline coupling x4, ConvAct(1->8->8->28)
conditioner, K=10 knots, batch 16384 per GPU), on N GPUs of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm: the oracle port on the host cores

One "step" = one pass of the hot path over one synthetic batch: NormalPrior draw with
log-density -> four checkerboard RQ-spline coupling steps (conditioner + spline +
log|det J|) -> phi^4 action, i.e. `model.posterior.sample__(B)`.  Ranks are independent
(batch sharding, no data-path collective): weak scaling, `value` = all ranks' samples
divided by the slowest rank's device time.

The JSON line carries, besides the base contract:
  roofline     : the dominant HBM-bound kernel (the RQ-spline apply), algorithmic bytes
                 ((8 + 4 P) B per site and step = 120 B at P = 28) over its mean
                 launch duration measured with CUDA events inside the timed steps, against
                 the measured copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline : the reference's operator sequence on torch CPU (oracle/torch_port.py, float64, all host
                 threads) timed on a bounded sample
  e2e          : the same metric through the public API with HOST buffers: the prior draw
                 comes from pinned host memory (H2D inside the timed region; the copy of step
                 i+1 runs on a second stream while step i is evaluated) and log q, log p are
                 read back (D2H) with a stream sync every step
"""

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LATTICE = (64, 64)
KNOTS = 10
HIDDEN = [8, 8]
N_STEPS_FLOW = 4
ACTION = dict(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5)
METRIC = "samples/sec (flow fwd+logJ+action) 64^2 phi^4"
WORKLOAD = ("configs[2]: 2-D scalar phi^4 64x64, RQ-spline coupling x4 (K=10, xlim=ylim=(-5,5), linear "
            "extrapolation), ConvAct(1->8->8->28, k=3, tanh, circular, no bias) conditioner")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="samples per GPU and step")
    ap.add_argument("--cpu-batch", type=int, default=256, help="samples per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------- CPU arm
def cpu_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def time_cpu_port(batch, steps, warmup):
    """The reference's CPU path for this workload: oracle/torch_port.py issues the same ATen
    operator sequence as the reference (float64, ATen threads = all host cores).  Test
    infrastructure used here only as the reported CPU baseline."""
    import torch
    from oracle import torch_port as T
    threads = cpu_threads()
    torch.set_num_threads(threads)
    gen = torch.Generator('cpu').manual_seed(0)
    sizes = [1] + HIDDEN + [3 * KNOTS - 2]
    nets = [[(torch.randn(sizes[i + 1], sizes[i], 3, 3, generator=gen, dtype=torch.float64, device='cpu')
              / math.sqrt(9 * sizes[i])) for i in range(3)] for _ in range(N_STEPS_FLOW)]
    mask = T.evenodd_mask(LATTICE)

    def one():
        x = torch.randn(batch, *LATTICE, generator=gen, dtype=torch.float64, device='cpu')
        with torch.no_grad():
            return T.posterior_sample__(x, nets, mask, (-5.0, 5.0), (-5.0, 5.0), ACTION)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt, threads


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is
    pure Python on torch and cannot travel to the GPU box, so this times its restatement on the
    same ATen operators (oracle/torch_port.py, float64, all host threads); each step is a bounded
    sample of `--cpu-batch` samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, dt, cores = time_cpu_port(args.cpu_batch, args.steps, max(args.warmup, 1))
    sample = (f"{args.cpu_batch} samples/step x {args.steps} steps of the same workload "
              "(torch ATen float64 port of the reference's operator sequence)")
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "lattice": list(LATTICE),
                   "cpu_sample_per_step": args.cpu_batch},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region."""

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------- GPU arm
def build_model(torch):
    from normflow__b200 import Model
    from normflow__b200.action import ScalarPhi4Action
    from normflow__b200.mask import EvenOddMask
    from normflow__b200.nn import ModuleList_, ConvAct, RQSplineCoupling_
    from normflow__b200.prior import NormalPrior
    torch.manual_seed(0)
    mask = EvenOddMask(shape=LATTICE)
    nets = [ConvAct(1, 3 * KNOTS - 2, 3, conv_dim=2, hidden_sizes=HIDDEN, acts=('tanh', 'tanh', None), bias=False)
            for _ in range(N_STEPS_FLOW)]
    net_ = ModuleList_([RQSplineCoupling_(nets, mask=mask, xlim=(-5, 5), ylim=(-5, 5),
                                          extrap=dict(left='linear', right='linear'))])
    model = Model(prior=NormalPrior(shape=LATTICE), net_=net_, action=ScalarPhi4Action(**ACTION))
    model.device_handler.to('cuda')
    return model


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary."""
    path = os.path.join(ROOT, "profiles", "dominant_kernel.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get("dram_bytes_per_launch")
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (normflow__b200 has no CPU path); "
                           "use --impl reference for the CPU arm")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from normflow__b200 import _C
    model = build_model(torch)
    B, V = args.batch, int(np.prod(LATTICE))
    torch.manual_seed(1234 + rank)          # independent Philox key per rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return model.posterior.sample__(B)

    # warm-up with exactly the timed loop's pattern (results of the previous step still alive
    # while the next one runs, per-kernel event spans on) so that the caching allocator and
    # the event pool are in steady state: a first-use cudaMalloc inside the timed region
    # stalls the host for ~15 ms and shows up as GPU idle time
    _C.kernel_timer = _C.KernelTimer()
    y = logq = logp = None
    for _ in range(max(args.warmup, 3)):
        y, logq, logp = step()
    barrier()

    # ---- timed region: K steps, device time, max over ranks -------------------------
    timer = _C.KernelTimer()
    _C.kernel_timer = timer
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.05)                           # let the sampler thread finish its NVML set-up
    n0 = _C.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        y, logq, logp = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _C.launch_count() - n0
    clock_info = clocks.finish()
    _C.kernel_timer = None
    kernels = timer.summary()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end with host buffers ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_x = torch.randn(B, *LATTICE, dtype=torch.float32, device="cpu").pin_memory()
        host_out = torch.empty(2, B, dtype=torch.float32, device="cpu").pin_memory()

        # The step's input arrives in host memory.  It is copied in chunks on a copy stream while
        # the compute stream evaluates the chunks already on the device (public API calls per
        # chunk: prior.log_prob, net_, action), so the PCIe transfer hides behind the kernels.
        # chunk boundaries: whole waves of the fused kernel's persistent grid (2 CTAs per SM), a short
        # first chunk so that compute starts early, then ~1/8 of the batch each
        wave = 2 * torch.cuda.get_device_properties(local).multi_processor_count
        bounds = [0]
        if B >= 16 * wave:
            bounds.append(2 * wave)
            step_c = max(wave, (B // 8) // wave * wave)
            while bounds[-1] + step_c < B:
                bounds.append(bounds[-1] + step_c)
        bounds.append(B)
        n_chunks = len(bounds) - 1
        # two device input buffers: while step i is evaluated, the copy stream already brings in step i + 1
        # (a loader with a prefetch depth of one); every step's copy is issued and completed inside the
        # timed region, only the copy of the first step cannot hide behind an earlier step
        x_dev = [torch.empty(B, *LATTICE, dtype=torch.float32, device="cuda") for _ in range(2)]
        res_dev = torch.empty(2, B, dtype=torch.float32, device="cuda")
        copy_stream = torch.cuda.Stream()
        ready = [[torch.cuda.Event() for _ in range(n_chunks)] for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def chunks_of(i):
            # only the first step's copy is exposed: it goes in chunks so that compute starts early; the later
            # ones complete behind the previous step's kernels and are evaluated whole
            return bounds if i == 0 else [0, B]

        def enqueue_copy(i):
            buf, bnd = i & 1, chunks_of(i)
            copy_stream.wait_event(consumed[buf])          # the step that last read this buffer is done with it
            with torch.cuda.stream(copy_stream):
                for c in range(len(bnd) - 1):
                    lo, hi = bnd[c], bnd[c + 1]
                    x_dev[buf][lo:hi].copy_(host_x[lo:hi], non_blocking=True)
                    ready[buf][c].record(copy_stream)

        def e2e_run(n_steps):
            main = torch.cuda.current_stream()
            for ev in consumed:
                ev.record(main)
            enqueue_copy(0)
            for i in range(n_steps):
                buf, bnd = i & 1, chunks_of(i)
                with torch.no_grad():
                    for c in range(len(bnd) - 1):
                        lo, hi = bnd[c], bnd[c + 1]
                        main.wait_event(ready[buf][c])
                        x = x_dev[buf][lo:hi]
                        logr = model.prior.log_prob(x)
                        yy, logJ = model.net_(x)
                        torch.sub(logr, logJ, out=res_dev[0, lo:hi])
                        torch.neg(model.action(yy), out=res_dev[1, lo:hi])
                    consumed[buf].record(main)
                    if i + 1 < n_steps:
                        enqueue_copy(i + 1)
                    host_out.copy_(res_dev, non_blocking=True)
                main.synchronize()                         # this step's result is in host memory

        e2e_run(2)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = t.item()
        e2e = {"value": world * B * args.steps / dt, "unit": "samples/s",
               "h2d_bytes_per_step": int(B * V * 4), "d2h_bytes_per_step": int(2 * B * 4),
               "chunks_first_step": n_chunks,
               "pipeline": "two device input buffers: the pinned-host batch of step i+1 is copied on a second stream "
                           "while step i is evaluated (all copies inside the timed region); per-step D2H + stream sync"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant HBM-bound kernel --------------------------------------
    pk, pk_src = peaks()
    P = 3 * KNOTS - 2
    # dominant kernel: the fused coupling step when the model takes that path, else the
    # unfused RQ-spline apply.  Both are scored under the SAME algorithmic byte model
    # (SURVEY 8d): one atomic step reads x, reads P conditioner channels, writes y.
    fused = kernels.get("fused2d_step", None)
    rq = fused or kernels.get("rqs_fwd", None)
    roofline = None
    if rq:
        bytes_per_launch = (8 + 4 * P) * V * B          # read x, read P params, write y
        achieved = bytes_per_launch / (rq["avg_ms"] * 1e-3) / 1e9
        name = ("nfk_fused2d_step (conditioner + RQ spline + log|det J| in one kernel)" if fused
                else "nfk_rqs_fwd (site_kernel<RqsOp<10,0>>)")
        roofline = {"kernel": name, "bound": "hbm", "achieved": achieved,
                    "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                    "traffic": ncu_traffic(), "peak_source": pk_src,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": rq["avg_ms"],
                    "share_of_step": rq["total_ms"] / ms}
        if fused:
            # the conditioner output of the byte model never reaches DRAM in the fused kernel
            # (real traffic = 8 B/site): what actually bounds it is the contraction
            flop = 2.0 * 9 * (HIDDEN[0] + HIDDEN[0] * HIDDEN[1] + HIDDEN[1] * P / 2) * V * B
            roofline["note"] = ("compute-bound kernel: 'achieved' is algorithmic bytes (incl. the never-materialised "
                                "conditioner output) per second; real DRAM traffic is 'traffic'")
            roofline["contraction_tflops"] = flop / (rq["avg_ms"] * 1e-3) / 1e12
    step_bytes = (4 + N_STEPS_FLOW * (8 + 4 * P) + 4) * V      # 488 V per sample
    total_ms = sum(k["total_ms"] for k in kernels.values())
    kernel_table = {name: {"launches": k["launches"], "avg_ms": round(k["avg_ms"], 4),
                           "share": round(k["total_ms"] / max(total_ms, 1e-9), 4)} for name, k in kernels.items()}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cb = args.cpu_batch
        v, dt, cores = time_cpu_port(cb, steps=4, warmup=1)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{cb} samples/step x 4 steps of the same workload, torch ATen float64 port of the "
                         f"reference's operator sequence, {dt:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "lattice": list(LATTICE),
                   "l2": "inputs larger than L2: every step draws a fresh 268 MB field batch (and writes 4 x 268 MB of intermediate fields), 126 MB L2",
                   "model_bytes_per_sample": step_bytes,
                   "hbm_model_frac_whole_step": step_bytes * value / world / 1e9 / pk["hbm_gbs"]},
        "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "kernels": kernel_table,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
