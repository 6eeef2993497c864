import sys
sys.path.insert(0, ".")
import numpy as np, scipy.sparse, torch
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR
dev = torch.device("cuda:0")
rng = np.random.default_rng(3)
items, f = 3000, 128
Y = rng.random((items, f)).astype(np.float32)
G64 = Y.astype(np.float64).T @ Y.astype(np.float64) + 0.1 * np.eye(f)
for n in (8, 16, 64):
    idx = np.sort(rng.choice(items, n, replace=False)).astype(np.int32)
    d = (10 * np.log1p(rng.integers(1, 6, size=n))).astype(np.float32)
    C = scipy.sparse.csr_matrix((d, idx, np.array([0, n], dtype=np.int64)), shape=(1, items))
    Yd = torch.from_numpy(Y).to(dev)
    X = engine.half_step(DeviceCSR.from_scipy(C, dev), Yd, engine.gram(Yd, 0.1), algo=_lib.ALGO_TCGEN05).cpu().numpy()[0]
    hdr = engine.workspace(0, dev)[0:16].view(torch.float32).cpu().numpy()
    print("n", n, "hdr maxdiagG", hdr[2], "maxd", hdr[3], "true", G64.diagonal().max(), d.max())
    Yr = Y[idx].astype(np.float64)
    W = (Yr.T * d) @ Yr
    b = ((d + 1)[:, None] * Yr).sum(0)
    rel = lambda x: np.linalg.norm(X - x) / np.linalg.norm(x)
    for wm in (0, 0.25, 0.5, 1, 2, 4):
        for bm in (0.5, 1, 2):
            print(f"   W x{wm:<5} b x{bm:<4} err {rel(np.linalg.solve(G64 + wm * W, bm * b)):.3e}")
    # b from a subset of the entries
    for k in (1, 2, 4):
        bs = ((d + 1)[:, None] * Yr)[::k].sum(0)
        print(f"   b from every {k}-th entry, W x1: err {rel(np.linalg.solve(G64 + W, bs)):.3e}")
