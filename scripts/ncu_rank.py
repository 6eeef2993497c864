"""Workload for ncu: config 5 (top-100 of all items) for two batches of users, tensor-core path."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from recmodel_b200 import engine
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
U = torch.from_numpy(rng.standard_normal((138493, 128)).astype(np.float32) * 0.1).to(dev)
V = torch.from_numpy(rng.standard_normal((26744, 128)).astype(np.float32) * 0.1).to(dev)
users = torch.arange(75776, device=dev, dtype=torch.int64)
for _ in range(2):
    ids = engine.score_topk(users, None, U, V, 100)
torch.cuda.synchronize()
print("ok", int(ids.sum()))
