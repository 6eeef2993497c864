"""Accuracy of the heaviest rows of the item half-step at ML-20M shape: tcgen05 and SIMT paths against an
fp64 solve of the same rows (the tensor core truncates its fp32 accumulation, scripts/probe/mma_round_probe.cu,
so long rows are where a bias would show)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from oracle import wmf_oracle as orc
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR
from recmodel_b200.synthetic import make_counts_cached
C = make_counts_cached(138493, 26744, 20_000_000)
C.data = orc.preprocess_counts(C.data)
dev = torch.device("cuda:0")
Cd = DeviceCSR.from_scipy(C, dev); CT = Cd.transpose()
Y = torch.from_numpy(orc.init_items(26744, 128, False)).to(dev)
U = engine.half_step(Cd, Y, engine.gram(Y, 0.1), algo=_lib.ALGO_TCGEN05)
G = engine.gram(U, 0.1)
Xt = engine.half_step(CT, U, G, algo=_lib.ALGO_TCGEN05).cpu().numpy()
Xs = engine.half_step(CT, U, G, algo=_lib.ALGO_SIMT).cpu().numpy()
CTh = C.T.tocsr()
Uh = U.cpu().numpy().astype(np.float64)
G64 = Uh.T @ Uh + 0.1 * np.eye(128)
lens = np.diff(CTh.indptr)
order = np.argsort(-lens)
rows = np.concatenate([order[:12], order[100:104], order[2000:2004]])
print("row  entries  tc-vs-fp64  simt-vs-fp64  ref32-vs-fp64")
for r in rows:
    lo, hi = CTh.indptr[r], CTh.indptr[r + 1]
    Yr = Uh[CTh.indices[lo:hi]]; d = CTh.data[lo:hi].astype(np.float64)
    x = np.linalg.solve(G64 + (Yr * d[:, None]).T @ Yr, (d + 1) @ Yr)
    Yr32 = Yr.astype(np.float32); d32 = CTh.data[lo:hi]
    G32 = (np.dot(Yr32.T * 0, Yr32) if False else None)
    x32 = np.linalg.solve(np.dot(Yr32.T, Yr32 * d32[:, None]) + G64.astype(np.float32), np.dot(d32 + 1, Yr32))
    e = lambda v: np.linalg.norm(v - x) / np.linalg.norm(x)
    print(f"{r:6d} {hi-lo:7d}  {e(Xt[r]):.2e}  {e(Xs[r]):.2e}  {e(x32):.2e}")
