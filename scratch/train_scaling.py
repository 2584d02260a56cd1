"""Weak scaling of the training step (Fitter.step incl. the flat gradient all-reduce): run under torchrun.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scratch/train_scaling.py [B_per_gpu]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = bench.build_model(torch)
model.device_handler.ddp_wrapper(rank, world, device=torch.device("cuda", local))
torch.manual_seed(100 + rank)
model.fit(n_epochs=2, batch_size=B, checkpoint_dict=dict(print_stride=1000, print_batch_size=64 * world, display=False))
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    model.fit.step()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# every rank must hold the same parameters after the averaged steps
flat = torch.cat([p.detach().flatten() for p in model.net_.parameters()])
ref = flat.clone()
if world > 1: dist.broadcast(ref, 0)
same = bool(torch.equal(flat, ref))
ok = torch.tensor([int(same)], device="cuda")
if world > 1: dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"train step: world={world} B/gpu={B}: {ms.item():.2f} ms/step -> {world * B / ms.item() * 1e3:.0f} samples/s; params identical on all ranks: {bool(ok.item())}")
if world > 1: dist.destroy_process_group()
