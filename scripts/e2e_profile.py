import sys, time; sys.path.insert(0, ".")
import numpy as np, torch
from recmodel_b200 import WMF, engine
from recmodel_b200.engine import DeviceCSR
from recmodel_b200.synthetic import make_counts_cached, split_train_test
C = make_counts_cached(138493, 26744, 20_000_000)
tr, te = split_train_test(C)
dev = torch.device("cuda:0")
def T(label, fn, n=3):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print(f"{label:34s} {min(ts)*1e3:8.2f} ms"); return r
m = WMF(num_items=26744, num_users=138493, dim=128, gamma=0.1, weighted=True)
T("train(iterations=1) total", lambda: m.train(tr, 1, eval_mat=te, count_mat=tr, cores=1, stopping_rounds=99))
print(m.last_train_stats)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); m.train(tr, 1, eval_mat=te, count_mat=tr, cores=1, stopping_rounds=99); u = m.users; pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
Cd = T("DeviceCSR.from_scipy(train)", lambda: DeviceCSR.from_scipy(tr, dev))
T("preprocess", lambda: engine.preprocess_(Cd.data.clone(), "log", 10, 1))
CT = T("transpose", lambda: Cd.transpose())
def sched(c):
    c._row_order = None; return c.row_order
T("schedule users", lambda: sched(Cd)); T("schedule items", lambda: sched(CT))
Ed = T("DeviceCSR.from_scipy(eval)", lambda: DeviceCSR.from_scipy(te, dev))
T("eval sddmm", lambda: engine.sddmm_loss(Ed, m.users_device, m.items_device))
T("factors D2H", lambda: (engine.d2h(m._users_d), engine.d2h(m._items_d)))
T("tr.tocsr()", lambda: tr.tocsr())
T("np.ascontiguousarray int64 indptr", lambda: np.ascontiguousarray(tr.indptr, dtype=np.int64))
