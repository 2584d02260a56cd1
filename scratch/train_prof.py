import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
B = int(sys.argv[1]); n = int(sys.argv[2])
model = bench.build_model(torch)
model.fit(n_epochs=1, batch_size=B, checkpoint_dict=dict(print_stride=1000, print_batch_size=64, display=False))
torch.cuda.synchronize()
for _ in range(n):
    model.fit.step()
torch.cuda.synchronize()
