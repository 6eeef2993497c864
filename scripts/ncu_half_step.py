"""Workload for ncu: 3 user half-steps then 3 item half-steps at ML-20M shape (tcgen05 kernel)."""
import sys
sys.path.insert(0, ".")
import torch
from oracle import wmf_oracle as orc  # initial factors only (R1 draw)
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR
from recmodel_b200.synthetic import make_counts_cached
C = make_counts_cached(138493, 26744, 20_000_000)
dev = torch.device("cuda:0")
Cd = DeviceCSR.from_scipy(C, dev); engine.preprocess_(Cd.data, "log", 10, 1); CT = Cd.transpose()
Y = torch.from_numpy(orc.init_items(26744, 128, False)).to(dev)
G = engine.gram(Y, 0.1)
for _ in range(3):
    U = engine.half_step(Cd, Y, G, algo=_lib.ALGO_TCGEN05)
G2 = engine.gram(U, 0.1)
for _ in range(3):
    V = engine.half_step(CT, U, G2, algo=_lib.ALGO_TCGEN05)
torch.cuda.synchronize()
print("ok", float(U.abs().mean()), float(V.abs().mean()))
