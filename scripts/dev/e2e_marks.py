"""Host-side time marks of WMF.train(iterations=1) at ML-20M shape, under torchrun or alone (development)."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import torch.distributed as dist
from recmodel_b200 import WMF
from recmodel_b200.synthetic import make_counts_cached, split_train_test
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    make_counts_cached(138493, 26744, 20_000_000)
if world > 1:
    dist.barrier()
C = make_counts_cached(138493, 26744, 20_000_000)
tr, te = split_train_test(C)
m = WMF(num_items=26744, num_users=138493, dim=128, gamma=0.1, weighted=True, device=torch.device("cuda", local))
for k in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.train(tr, 1, eval_mat=te, count_mat=tr, cores=1, stopping_rounds=99)
    t1 = time.perf_counter()
    u, i = m.users, m.items
    t2 = time.perf_counter()
    if k == 3:
        print(f"rank {rank}: train {1e3 * (t1 - t0):.2f} ms, read-back {1e3 * (t2 - t1):.2f} ms; marks {m.last_train_stats.get('host_marks_ms')} "
              f"half-steps {m.last_train_stats['half_step_ms']} eval {m.last_train_stats['eval_ms']}", flush=True)
if world > 1:
    dist.barrier()
    os._exit(0)
