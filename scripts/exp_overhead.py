import os, sys
sys.path.insert(0, ".")
import torch
from oracle import wmf_oracle as orc
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR, device_schedule
from recmodel_b200.synthetic import make_counts_cached
C = make_counts_cached(138493, 26744, 20_000_000)
dev = torch.device("cuda:0")
Cd = DeviceCSR.from_scipy(C, dev); engine.preprocess_(Cd.data, "log", 10, 1); CT = Cd.transpose()
Y = torch.from_numpy(orc.init_items(26744, 128, False)).to(dev)
G = engine.gram(Y, 0.1)
def t(csr, Yd, G):
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); X = engine.half_step(csr, Yd, G, algo=_lib.ALGO_TCGEN05); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, X
for ov in (4, 8, 12, 16, 24):
    Cd._row_order = device_schedule(Cd.indptr, 148, row_overhead=ov)
    CT._row_order = device_schedule(CT.indptr, 148, row_overhead=ov)
    tu, U = t(Cd, Y, G)
    G2 = engine.gram(U, 0.1)
    ti, _ = t(CT, U, G2)
    print(f"row_overhead {ov}: user {tu:.2f} ms  item {ti:.2f} ms  (sched len {Cd._row_order.numel()}, {CT._row_order.numel()})", flush=True)
