"""Run under torchrun on N GPUs: the row-sharded WMF.train must reproduce the single-GPU
factors bit for bit (row -> rank assignment never changes a row's arithmetic)."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch, torch.distributed as dist
from recmodel_b200 import WMF
from recmodel_b200.synthetic import make_counts, split_train_test

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
full = make_counts(20000, 6000, 1_500_000, seed=77)
tr, te = split_train_test(full)
res = {}
for dim, bias in ((128, False), (32, True)):
    m = WMF(num_items=6000, num_users=20000, dim=dim, gamma=0.1, weighted=True, bias=bias)
    it = m.train(tr, 2, eval_mat=te, count_mat=tr, cores=1, stopping_rounds=99)
    mse = float(m.eval_prec(te))
    # user-sharded ranking: every rank scores a slice of the users, the id lists are all-gathered
    torch.cuda.synchronize(); t0 = time.perf_counter()
    top = m.rank_batch(np.arange(6000), np.arange(20000), 50, distributed=True)  # collective
    torch.cuda.synchronize(); t_rank = time.perf_counter() - t0
    res[(dim, bias)] = (m.users.copy(), m.items.copy(), mse, it, top, t_rank)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    # single-GPU run of the same thing in this process (no process group -> unsharded)
    for (dim, bias), (U, V, mse, it, top, t_rank) in res.items():
        m = WMF(num_items=6000, num_users=20000, dim=dim, gamma=0.1, weighted=True, bias=bias)
        it1 = m.train(tr, 2, eval_mat=te, count_mat=tr, cores=1, stopping_rounds=99)
        same_u, same_v = np.array_equal(m.users, U), np.array_equal(m.items, V)
        print(f"dim={dim} bias={bias} world={world}: users bitwise equal={same_u} items bitwise equal={same_v} "
              f"mse sharded={mse:.6f} single={float(m.eval_prec(te)):.6f} iter {it}=={it1}")
        assert same_u and same_v and it == it1
        torch.cuda.synchronize(); t0 = time.perf_counter()
        top1 = m.rank_batch(np.arange(6000), np.arange(20000), 50)
        torch.cuda.synchronize(); t1 = time.perf_counter() - t0
        print(f"   rank_batch top-50 of 6000 items for 20000 users: sharded lists identical={np.array_equal(top, top1)} "
              f"({t_rank * 1e3:.2f} ms on {world} GPUs, {t1 * 1e3:.2f} ms on one)")
        assert np.array_equal(top, top1)
        assert abs(mse - float(m.eval_prec(te))) < 1e-6 * abs(mse)
    print("multi-gpu ok")
