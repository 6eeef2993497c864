"""Small-case localisation of half-step errors: rows of chosen lengths, tcgen05 vs fp64."""
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


def case(name, lens, use_order=True):
    rows = len(lens)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([np.sort(rng.choice(items, n, replace=False)) for n in lens]).astype(np.int32) if sum(lens) else np.zeros(0, np.int32)
    d = (10 * np.log1p(rng.integers(1, 6, size=idx.size))).astype(np.float32)
    C = scipy.sparse.csr_matrix((d, idx, indptr), shape=(rows, items))
    Yd = torch.from_numpy(Y).to(dev)
    Gd = engine.gram(Yd, 0.1)
    X = engine.half_step(DeviceCSR.from_scipy(C, dev), Yd, Gd, algo=_lib.ALGO_TCGEN05, use_row_order=use_order)
    torch.cuda.synchronize()
    flags = engine.workspace(0, dev)[0:8].view(torch.int32).cpu().numpy()
    X = X.cpu().numpy()
    errs = []
    for r in range(rows):
        sl = slice(indptr[r], indptr[r + 1])
        Yr = Y[idx[sl]].astype(np.float64)
        A = G64 + (Yr.T * d[sl]) @ Yr
        b = ((d[sl] + 1)[:, None] * Yr).sum(0)
        x = np.linalg.solve(A, b) if lens[r] else np.zeros(f)
        errs.append(np.linalg.norm(X[r] - x) / max(np.linalg.norm(x), 1e-30))
    errs = np.array(errs)
    print(f"{name:34s} rows={rows:5d} flags={flags} max err {errs.max():.2e} median {np.median(errs):.2e} bad rows {(errs > 1e-4).sum()}", flush=True)
    return errs


case("1 row n=8", [8])
case("1 row n=16", [16])
case("1 row n=20", [20])
case("1 row n=32", [32])
case("1 row n=40", [40])
case("1 row n=64", [64])
case("1 row n=200", [200])
case("1 row n=5000", [2500])
case("5 rows n=8", [8] * 5)
case("600 rows n=8", [8] * 600)
case("600 rows n=40", [40] * 600)
case("600 rows n=100", [100] * 600)
case("2000 rows mixed", list(rng.integers(0, 300, size=2000)))
