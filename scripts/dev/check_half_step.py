"""Development check of the tcgen05 half-step pipeline (whitening + dual + primal + fix-up) against the fp64
restatement, with per-row-length error buckets. Usage (GPU box):
    WMF_TC_DUAL=0|1 python scripts/dev/check_half_step.py [case ...]
"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, scipy.sparse
from oracle import wmf_oracle as orc
from recmodel_b200 import _lib, engine
from recmodel_b200.engine import DeviceCSR
from recmodel_b200.synthetic import make_counts

dev = torch.device("cuda:0")
CASES = {
    # name: users, items, nnz, f, bias
    "f128": (3000, 2000, 300_000, 128, False),
    "f128_short": (4000, 3000, 120_000, 128, False),
    "f64": (2000, 1500, 150_000, 64, False),
    "f65b": (2000, 1500, 150_000, 65, True),
    "f129b": (1500, 1200, 90_000, 129, True),
    "f256": (1500, 1200, 60_000, 256, False),
    "f16": (500, 300, 9_000, 16, False),
    "f192": (1500, 1200, 120_000, 192, False),
    "f256_long": (600, 30_000, 400_000, 256, False),   # long rows: segments of the 256-wide kernel
}


def run(name):
    users, items, nnz, f, bias = CASES[name]
    C = make_counts(users, items, nnz, seed=31)
    C.data = orc.preprocess_counts(C.data)
    rng = np.random.default_rng(3)
    Y = rng.random((items, f)).astype(np.float32)
    Yd = torch.from_numpy(Y).to(dev)
    Cd = DeviceCSR.from_scipy(C, dev)
    G = engine.gram(Yd, 0.1, ones_col0=bias)
    t0 = time.time()
    X = engine.half_step(Cd, Yd, G, bias=bias, algo=_lib.ALGO_TCGEN05)
    torch.cuda.synchronize()
    flags, fixed = engine.half_step_status()
    X = X.cpu().numpy()
    step = orc.half_step_bias if bias else orc.half_step
    rows = np.arange(min(users, 1200))
    x64 = step(Y, C[rows], 0.1, np.float64)
    lens = np.diff(C.indptr)[rows]
    err = np.linalg.norm(X[rows] - x64, axis=1) / np.maximum(np.linalg.norm(x64, axis=1), 1e-30)
    err[lens == 0] = np.abs(X[rows][lens == 0]).max(axis=1) if (lens == 0).any() else 0
    print(f"== {name}: {users}x{items} nnz {C.nnz} f {f} bias {bias}  dual_max {_lib.load().wmf_als_dual_max_entries()} "
          f"WMF_TC_DUAL={os.environ.get('WMF_TC_DUAL', '1')} flags {flags} fixup rows {fixed} finite {np.isfinite(X).all()} "
          f"({time.time() - t0:.2f}s)")
    for lo, hi in ((0, 0), (1, 8), (9, 16), (17, 32), (33, 64), (65, 96), (97, 128), (129, 256), (257, 10**9)):
        sel = (lens >= lo) & (lens <= hi)
        if sel.any():
            e = err[sel]
            print(f"   rows with {lo:>4}..{hi if hi < 10**9 else 'inf':>4} entries: {sel.sum():5d}  max err {np.nanmax(e):.2e}  median {np.nanmedian(e):.2e}  nan {np.isnan(e).sum()}")
    worst = np.argsort(-np.nan_to_num(err, nan=1e9))[:3]
    for r in worst:
        print(f"   worst row {rows[r]} n={lens[r]} err {err[r]:.3e}")
    return float(np.nanmax(err)) if np.isfinite(err).all() else float("inf")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    bad = 0
    for n in names:
        e = run(n)
        bad += e > 1e-4
    print("dev check", "FAILED" if bad else "ok")
    sys.exit(1 if bad else 0)
