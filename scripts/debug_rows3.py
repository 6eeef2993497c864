import os, sys
sys.path.insert(0, ".")
import numpy as np, scipy.sparse, torch
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR
dev = torch.device("cuda:0")
rng = np.random.default_rng(3)
items, f = 3000, 128
Y = rng.random((items, f)).astype(np.float32)
G64 = Y.astype(np.float64).T @ Y.astype(np.float64) + 0.1 * np.eye(f)
n = int(sys.argv[1])
mode = int(os.environ.get("WMF_TC_DEBUG", "0"))
idx = np.sort(rng.choice(items, n, replace=False)).astype(np.int32)
d = (10 * np.log1p(rng.integers(1, 6, size=n))).astype(np.float32)
C = scipy.sparse.csr_matrix((d, idx, np.array([0, n], dtype=np.int64)), shape=(1, items))
Yd = torch.from_numpy(Y).to(dev)
X = engine.half_step(DeviceCSR.from_scipy(C, dev), Yd, engine.gram(Yd, 0.1), algo=_lib.ALGO_TCGEN05).cpu().numpy()[0]
Yr = Y[idx].astype(np.float64)
W = (Yr.T * d) @ Yr
b = ((d + 1)[:, None] * Yr).sum(0)
np.set_printoptions(precision=4, linewidth=200, suppress=True)
if mode == 1:
    print("rhs: max rel err", np.max(np.abs(X - b) / np.abs(b)))
    print(X[:8]); print(b[:8])
elif mode >= 2:
    c = mode - 2
    ref = (G64 + W)[:, c]
    print(f"column {c} of A: max rel err", np.max(np.abs(X - ref) / np.abs(ref)))
    print("gpu  ", X[:10]); print("ref  ", ref[:10]); print("G    ", G64[:10, c]); print("W    ", W[:10, c])
    print("(gpu-G)/W", ((X - G64[:, c]) / W[:, c])[:16])
