import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from recmodel_b200 import engine
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
U = torch.from_numpy(rng.standard_normal((138493, 128)).astype(np.float32) * 0.1).to(dev)
V = torch.from_numpy(rng.standard_normal((26744, 128)).astype(np.float32) * 0.1).to(dev)
users = torch.arange(138493, device=dev, dtype=torch.int64)
engine.score_topk(users, None, U, V, 100); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    engine.score_topk(users, None, U, V, 100); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=60))
