"""Config 5: top-100 over all items for all users at ML-20M shape (exact fp32 scores)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from recmodel_b200 import engine
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
U = torch.from_numpy(rng.standard_normal((138493, 128)).astype(np.float32) * 0.1).to(dev)
V = torch.from_numpy(rng.standard_normal((26744, 128)).astype(np.float32) * 0.1).to(dev)
for nu in (4096, 138493):
    users = torch.arange(nu, device=dev, dtype=torch.int64)
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        ids = engine.score_topk(users, None, U, V, 100)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"users {nu}: {dt*1e3:.1f} ms  {nu/dt:.0f} users/s  {2*nu*26744*128/dt/1e12:.1f} TFLOP/s-equivalent", flush=True)
# exactness: tensor-core path vs the exact CUDA-core kernels (same ids, same order)
import os
users = torch.arange(0, 138493, 37, device=dev, dtype=torch.int64)
a = engine.score_topk(users, None, U, V, 100).cpu().numpy()
os.environ["WMF_SCORE_EXACT"] = "1"
b = engine.score_topk(users, None, U, V, 100).cpu().numpy()
del os.environ["WMF_SCORE_EXACT"]
print("tc vs exact: identical", bool((a == b).all()), "users", len(users), "mismatching rows", int((a != b).any(axis=1).sum()))
cand = torch.from_numpy(rng.permutation(26744)[:9000].astype(np.int64)).to(dev)
a = engine.score_topk(users[:500], cand, U, V, 50).cpu().numpy()
os.environ["WMF_SCORE_EXACT"] = "1"
b = engine.score_topk(users[:500], cand, U, V, 50).cpu().numpy()
print("candidate list: identical", bool((a == b).all()))
