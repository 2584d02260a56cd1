"""Functions run in spawned ranks (they have to be importable by the children)."""
import os

import torch


def fit_and_dump(model, out_dir, n_epochs, batch_size):
    """model.fit on this rank's share of the batch, then save the parameters this rank ended with."""
    h = model.device_handler
    model.fit(n_epochs=n_epochs, batch_size=batch_size // h.nranks,
              checkpoint_dict=dict(print_stride=1000, print_batch_size=64, display=False))
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten().cpu() for p in model.net_.parameters()])
    torch.save({"params": flat, "loss": list(model.fit.train_history['loss']), "device": str(flat.device),
                "rank": h.rank, "cuda": torch.cuda.current_device()}, os.path.join(out_dir, f"rank{h.rank}.pt"))


def fit_graph_and_dump(model, out_dir, n_epochs, batch_size):
    """The same with the optimisation step replayed as ONE captured CUDA graph that contains the NCCL
    all-reduce of the flat gradient buffer (Fitter.cuda_graph with nranks > 1)."""
    model.fit.cuda_graph = True
    fit_and_dump(model, out_dir, n_epochs, batch_size)
