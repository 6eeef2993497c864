import os, sys
sys.path.insert(0, ".")
import torch
from oracle import wmf_oracle as orc
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR
from recmodel_b200.synthetic import make_counts_cached
C = make_counts_cached(138493, 26744, 20_000_000)
dev = torch.device("cuda:0")
Cd = DeviceCSR.from_scipy(C, dev); engine.preprocess_(Cd.data, "log", 10, 1); CT = Cd.transpose()
Y = torch.from_numpy(orc.init_items(26744, 128, False)).to(dev)
G = engine.gram(Y, 0.1)
def t(csr, Yd, G):
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); X = engine.half_step(csr, Yd, G, algo=_lib.ALGO_TCGEN05); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), X
for ng in (4, 2, 1):
    os.environ["WMF_TC_GROUPS"] = str(ng)
    tu, U = t(Cd, Y, G)
    G2 = engine.gram(U, 0.1)
    ti, _ = t(CT, U, G2)
    print(f"groups {ng}: user {tu:.2f} ms  item {ti:.2f} ms", flush=True)
