"""Seeded synthetic implicit-feedback matrices.

The reference ships no data (its notebook loads a Netflix.npz that is not in the
repository), so every parity test and bench line in this repo runs on matrices from
this generator: distinct (user, item) pairs, item popularity ~ rank^-0.8, log-normal
user activity, integer counts 1..5 stored as float32, indices sorted per row. An
optional planted low-rank structure makes Recall@N meaningfully above chance.

Shapes of BASELINE.json's configs are in ``SHAPES``.
"""
import numpy as np
import scipy.sparse

DEFAULT_SEED = 20240229

SHAPES = {
    # name: (users, items, nnz, dim, bias)
    "cfg1": (6040, 3706, 1_000_000, 64, True),      # ML-1M shape, CPU-reference config
    "cfg2": (138_493, 26_744, 20_000_000, 128, False),  # ML-20M shape, 1 B200
    "cfg3": (480_189, 17_770, 100_000_000, 128, False),  # Netflix shape
    "cfg4": (10_000_000, 1_000_000, 1_000_000_000, 256, False),
    "cfg4_scaled": (1_000_000, 100_000, 100_000_000, 256, False),   # config 4 at 1/10 of every dimension (one GPU)
}


def _popularity(n_items, exponent, rng):
    w = np.arange(1, n_items + 1, dtype=np.float64) ** (-exponent)
    rng.shuffle(w)  # popular items are not the low ids
    return w / w.sum()


def make_counts_cached(n_users, n_items, nnz, seed=DEFAULT_SEED, planted_rank=0, cache_dir=None):
    """``make_counts`` with an on-disk cache (generation of 20 M entries takes ~40 s of host
    time; tests and the bench in one session share it)."""
    import os
    cache_dir = cache_dir or os.environ.get("WMF_SYNTH_CACHE", "/tmp/wmf_synth_cache")
    path = os.path.join(cache_dir, f"counts_{n_users}_{n_items}_{nnz}_{seed}_{planted_rank}.npz")
    if os.path.exists(path):
        try:
            return scipy.sparse.load_npz(path)
        except Exception:
            pass
    mat = make_counts(n_users, n_items, nnz, seed=seed, planted_rank=planted_rank)
    try:
        os.makedirs(cache_dir, exist_ok=True)
        scipy.sparse.save_npz(path + ".tmp.npz", mat, compressed=False)
        os.replace(path + ".tmp.npz", path)
    except OSError:
        pass
    return mat


def make_counts(n_users, n_items, nnz, seed=DEFAULT_SEED, planted_rank=0, exponent=0.8,
                sigma=1.0, max_row_frac=0.5):
    """CSR float32 count matrix with exactly ``nnz`` distinct entries (or slightly fewer if
    the shape cannot hold them). Deterministic in ``seed``."""
    rng = np.random.default_rng(seed)
    nnz = int(min(nnz, int(n_users * n_items * max_row_frac)))
    act = rng.lognormal(mean=0.0, sigma=sigma, size=n_users)
    act /= act.sum()
    pop = _popularity(n_items, exponent, rng)
    cum_u = np.cumsum(act)
    cum_u[-1] = 1.0
    cum_i = np.cumsum(pop)
    cum_i[-1] = 1.0

    if planted_rank > 0:
        # users and items get a cluster; a user draws most items from its own cluster
        ucl = rng.integers(0, planted_rank, size=n_users)
        icl = rng.integers(0, planted_rank, size=n_items)
        order = np.argsort(icl, kind="stable")
        starts = np.searchsorted(icl[order], np.arange(planted_rank + 1))

    keys = np.empty(0, dtype=np.int64)
    want = nnz
    rounds = 0
    while keys.size < nnz and rounds < 30:
        m = int((want - keys.size) * 1.25) + 1024
        u = np.searchsorted(cum_u, rng.random(m), side="right").astype(np.int64)
        i = np.searchsorted(cum_i, rng.random(m), side="right").astype(np.int64)
        if planted_rank > 0:
            own = rng.random(m) < 0.8
            c = ucl[u]
            lo = starts[c]
            span = np.maximum(starts[c + 1] - lo, 1)
            inside = order[np.minimum(lo + (rng.random(m) * span).astype(np.int64), n_items - 1)]
            i = np.where(own, inside, i)
        np.minimum(u, n_users - 1, out=u)
        np.minimum(i, n_items - 1, out=i)
        keys = np.unique(np.concatenate([keys, u * n_items + i]))
        rounds += 1
    if keys.size > nnz:
        drop = rng.choice(keys.size, size=keys.size - nnz, replace=False)
        mask = np.ones(keys.size, dtype=bool)
        mask[drop] = False
        keys = keys[mask]
    rows = (keys // n_items).astype(np.int64)
    cols = (keys % n_items).astype(np.int32)
    vals = rng.integers(1, 6, size=keys.size).astype(np.float32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n_users), out=indptr[1:])
    idx_dtype = np.int32 if keys.size < 2**31 - 1 else np.int64
    mat = scipy.sparse.csr_matrix((vals, cols, indptr.astype(idx_dtype)), shape=(n_users, n_items))
    mat.has_sorted_indices = True  # keys are sorted, so columns ascend inside each row
    return mat


def make_counts_device(n_users, n_items, nnz, device, seed=DEFAULT_SEED, exponent=0.8, sigma=1.0):
    """``make_counts`` with the same recipe generated ON THE DEVICE (log-normal user activity, item popularity
    ~ rank^-exponent, distinct pairs, counts 1..5), for the shapes whose host generation takes minutes (Netflix
    shape, the 1 B-entry power-law config and its scaled version). Returns (indptr int64, indices int32, data
    float32) CUDA tensors of a CSR matrix with sorted indices; deterministic in ``seed`` on a given GPU model.
    torch is used for data generation only (not a product path)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    act = torch.exp(sigma * torch.randn(n_users, device=device, generator=g, dtype=torch.float64))
    cum_u = torch.cumsum(act / act.sum(), 0)
    pop = torch.arange(1, n_items + 1, device=device, dtype=torch.float64) ** -exponent
    pop = pop[torch.randperm(n_items, device=device, generator=g)]
    cum_i = torch.cumsum(pop / pop.sum(), 0)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    rounds = 0
    while keys.numel() < nnz and rounds < 40:
        m = int((nnz - keys.numel()) * 1.3) + 1024
        u = torch.searchsorted(cum_u, torch.rand(m, device=device, generator=g, dtype=torch.float64)).clamp_(max=n_users - 1)
        i = torch.searchsorted(cum_i, torch.rand(m, device=device, generator=g, dtype=torch.float64)).clamp_(max=n_items - 1)
        keys = torch.unique(torch.cat([keys, u * n_items + i]))
        del u, i
        rounds += 1
    if keys.numel() > nnz:
        keys = keys[torch.randperm(keys.numel(), device=device, generator=g)[:nnz]].sort().values
    rows = keys // n_items
    cols = (keys % n_items).to(torch.int32)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n_users), 0)
    data = torch.randint(1, 6, (keys.numel(),), device=device, generator=g).to(torch.float32)
    return indptr, cols, data


def split_train_test(matrix, train=0.8, seed=1993):
    """Copy-safe restatement of the reference's split (/root/reference/RecModel/utils.py:19-35):
    ``np.random.seed(seed)``; entry k (CSR data order) goes to train iff ``rand(nnz)[k] < train``.
    The reference aliases both halves under SciPy >= 1.15 and returns
    empty matrices; this version keeps its RNG semantics and fixes the aliasing."""
    matrix = matrix.tocsr()
    np.random.seed(seed)
    is_train = np.random.rand(len(matrix.data)) < train
    out = []
    for keep in (is_train, ~is_train):
        coo = matrix.tocoo(copy=True)
        coo.data = np.where(keep, coo.data, 0).astype(matrix.dtype)
        part = coo.tocsr()
        part.eliminate_zeros()
        out.append(part)
    return out
