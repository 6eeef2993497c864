"""CPU oracle for the WMF train + top-N path. TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product path (recmodel_b200/) never does, and raises if the CUDA
library is missing instead of falling back to anything here.

What this restates: the NumPy/SciPy/LAPACK arithmetic of titoeb/RecModel's WMF model, cited
per function as /root/reference/<file>:<lines>. The arithmetic itself lives in third-party
libraries the reference does not pin (README.md:22-33 names NumPy/SciPy only): np.dot (BLAS
sgemm/sgemv), np.linalg.solve (LAPACK sgesv: LU with partial pivoting), np.linalg.inv,
NumPy's pairwise add.reduce, np.argpartition / np.argsort. The oracle calls the same NumPy
entry points in the same order and dtype, so on one machine it reproduces the reference to
rounding of the BLAS backend.

Parity pinning: the reference has NO tests or golden vectors of its own. The
oracle is therefore pinned against the reference itself, executed unmodified in the build
container through oracle/ref_shim.py; the outputs are committed as tests/golden/*.npz by
tests/golden/make_golden.py and re-checked by tests/test_oracle_golden.py.

Every function takes ``dtype``: np.float32 reproduces the reference's default arithmetic;
np.float64 is the high-precision restatement used to bound the fp32 noise floor.
"""
import numpy as np
import scipy.sparse


# --------------------------------------------------------------------------------------
# R1  WMF.__init__                                  /root/reference/RecModel/wmf_model.py:10-23
# --------------------------------------------------------------------------------------
def init_items(num_items, dim, bias, seed=1993, dtype=np.float32):
    """Seeds the GLOBAL NumPy RNG and draws U[0,1) float64, then casts (wmf_model.py:11-17).
    With bias the factor matrix has dim+1 columns and column 0 is the item bias."""
    np.random.seed(seed)
    cols = dim + 1 if bias else dim
    return np.random.random((num_items, cols)).astype(dtype)


# --------------------------------------------------------------------------------------
# R3  count preprocessing                           wmf_model.py:119-126 (and :65-70)
# --------------------------------------------------------------------------------------
def preprocess_counts(data, mode="log", alpha=10, beta=1):
    """d = alpha*log(1+beta*x) | alpha*x, in the dtype of ``data``; ValueError otherwise."""
    if mode == "log":
        return alpha * np.log(1 + beta * data)
    if mode == "linear":
        return alpha * data
    raise ValueError(f"Pre_process_count {mode} is not implement please use log or linear.")


# --------------------------------------------------------------------------------------
# R4 / R6  recompute_factors (+ _par/_intern)       wmf_model.py:213-240, :242-250, :289-309
# --------------------------------------------------------------------------------------
def half_step(Y, C, lam, dtype=np.float32):
    """One ALS half-step without biases.

    x_r = (Y^T Y + lam I + sum_j d_j y_j y_j^T)^-1  sum_j (d_j + 1) y_j   over the stored
    entries j of row r of C (d = confidence minus one); rows with no entries give 0
    (wmf_model.py:223-225). The Pool variant (:242-250) computes the same per-row values."""
    Y = np.asarray(Y, dtype=dtype)
    f = Y.shape[1]
    G = np.dot(Y.T, Y) + lam * np.eye(f, dtype=dtype)
    out = np.empty((C.shape[0], f), dtype=dtype)
    ptr, ind, val = C.indptr, C.indices, C.data
    for r in range(C.shape[0]):
        lo, hi = ptr[r], ptr[r + 1]
        if lo == hi:
            out[r] = 0
            continue
        d = val[lo:hi].astype(dtype, copy=False)
        Yr = Y[ind[lo:hi]]
        A = np.dot(Yr.T, Yr * d[:, None]) + G
        out[r] = np.linalg.solve(A, np.dot(d + 1, Yr))
    return out


# --------------------------------------------------------------------------------------
# R5 / R6  recompute_factors_bias (+ _par/_intern)  wmf_model.py:311-351, :252-287
# --------------------------------------------------------------------------------------
def half_step_bias(Y, C, lam, dtype=np.float32):
    """One ALS half-step with biases, exactly the reference's formula (not the textbook one):
    beta = Y[:,0]; Y~ = Y with column 0 := 1; G = Y~^T Y~ + lam I (ones column regularised too,
    :332); per row d~ = d - beta[idx] (:343) and x_r = (G + Y~r^T diag(d~) Y~r)^-1 (d~+1) Y~r.
    Column 0 of the result is the row side's new bias. Does NOT mutate its argument (the
    reference does, :331; train() shields that with .copy(), :151-156). The serial version has
    no empty-row branch (solve(G, 0) = 0); the Pool version returns zeros (:274-276)."""
    Yt = np.array(Y, dtype=dtype, copy=True)
    beta = Yt[:, 0].copy()
    Yt[:, 0] = 1
    f = Yt.shape[1]
    G = np.dot(Yt.T, Yt) + lam * np.eye(f, dtype=dtype)
    out = np.empty((C.shape[0], f), dtype=dtype)
    ptr, ind, val = C.indptr, C.indices, C.data
    for r in range(C.shape[0]):
        lo, hi = ptr[r], ptr[r + 1]
        idx = ind[lo:hi]
        d = val[lo:hi].astype(dtype, copy=False) - beta[idx]
        Yr = Yt[idx]
        A = np.dot(Yr.T, Yr * d[:, None]) + G
        out[r] = np.linalg.solve(A, np.dot(d + 1, Yr))
    return out


# --------------------------------------------------------------------------------------
# R7  unweighted half-steps                         wmf_model.py:85, :88
# --------------------------------------------------------------------------------------
def unweighted_half_step(Y, R, lam, dim, dtype=np.float32):
    """X = (inv(Y^T Y + lam I_dim) Y^T R^T)^T for the fixed side Y and ratings R [rows x N].
    Uses np.eye(dim) like the reference, so bias=True (dim+1 columns) raises a shape error
    exactly as wmf_model.py:85 does."""
    Y = np.asarray(Y, dtype=dtype)
    W = np.dot(np.linalg.inv(np.dot(Y.T, Y) + lam * np.eye(dim, dtype=dtype)), Y.T)
    return scipy.sparse.csr_matrix.dot(W, R.T).T.copy()


# --------------------------------------------------------------------------------------
# R8  WMF.predict                                   wmf_model.py:191-211
# --------------------------------------------------------------------------------------
def pairwise_sum_rows(P):
    """Row sums of a 2-D float array in NumPy's pairwise add.reduce order, restated
    explicitly (numpy/core/src/umath/loops_utils.h.src pairwise_sum, unpinned by the
    reference): n < 8 sequential starting from the first element; n <= 128 eight strided
    accumulators over blocks of 8, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the
    scalar tail; n > 128 split at n/2 rounded down to a multiple of 8, left + right."""
    P = np.ascontiguousarray(P)
    n = P.shape[1]
    if n < 8:
        acc = P[:, 0].copy() if n else np.zeros(P.shape[0], P.dtype)
        for i in range(1, n):
            acc = acc + P[:, i]
        return acc
    if n <= 128:
        r = [P[:, k].copy() for k in range(8)]
        i = 8
        while i + 8 <= n:
            for k in range(8):
                r[k] = r[k] + P[:, i + k]
            i += 8
        acc = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            acc = acc + P[:, i]
            i += 1
        return acc
    half = (n // 2) - ((n // 2) % 8)
    return pairwise_sum_rows(P[:, :half]) + pairwise_sum_rows(P[:, half:])


def predict(user_f, item_f, users, items, bias):
    """Element-wise pair scores in the reference's rounding order: products are rounded to the
    factor dtype first, then summed by NumPy's pairwise reduce; with bias the latent sum, then
    + user bias, then + item bias (wmf_model.py:206, :209-211). A scalar / length-1 user
    broadcasts against many items. Length check as in :200-203."""
    if isinstance(users, (list, np.ndarray)) and isinstance(items, (list, np.ndarray)):
        if len(users) != len(items) and not (len(users) == 1 or len(items) == 0):
            raise ValueError("users and items need to have the same length or only one "
                             "user / item needs to be provided.")
    if not bias:
        return (user_f[users, :] * item_f[items, :]).sum(axis=1)
    ub = user_f[:, 0][users]
    ib = item_f[:, 0][items]
    return (user_f[:, 1:][users, :] * item_f[:, 1:][items, :]).sum(axis=1) + ub + ib


# --------------------------------------------------------------------------------------
# R10  WMF.rank                                     wmf_model.py:25-47
# --------------------------------------------------------------------------------------
def rank(user_f, item_f, items, user, topn=None, bias=False):
    """Top-``topn`` candidate ids for one user, best first. Same two branches as the reference
    (:40-47): argpartition when topn < len/2, else a full argsort. Tie order is unspecified in
    both, so parity is on index sets / order up to equal scores."""
    items = np.asarray(items)
    if topn is None:
        topn = len(items)
    s = predict(user_f, item_f, user, items, bias)
    if len(s) * 0.5 > topn:
        sel = np.argpartition(s, list(range(-topn, 0, 1)))[-topn:]
    else:
        sel = np.argsort(s)[-topn:]
    return items[sel][::-1]


def rank_scores(user_f, item_f, items, user, bias=False):
    """The score vector rank() orders by (helper for tie-aware comparisons in tests)."""
    return predict(user_f, item_f, user, np.asarray(items), bias)


# --------------------------------------------------------------------------------------
# R9  RecModel.eval_prec                            base_model.py:150-179
# --------------------------------------------------------------------------------------
def eval_prec(user_f, item_f, mat, bias, metric="mse"):
    """mean((r_ui - yhat_ui)^2) (or RMSE / MAE) over entries of ``mat`` whose stored value is
    non-zero (``nonzero()`` drops explicit zeros, base_model.py:163)."""
    metric = metric.upper()
    if metric not in ("MSE", "RMSE", "MAE"):
        raise ValueError("Metric {metric} is not implemented.")
    rows, cols = mat.nonzero()
    pred = predict(user_f, item_f, rows, cols, bias).reshape(1, -1)
    diff = mat[rows, cols] - pred
    if metric == "MAE":
        return np.mean(np.abs(diff))
    mse = np.mean(np.square(diff))
    return np.sqrt(mse) if metric == "RMSE" else mse


def eval_prec_f64(user_f, item_f, mat, bias, metric="mse"):
    """Same quantity accumulated in float64 (predictions still rounded in the factor dtype);
    the comparison target for the device reduction, which sums in double."""
    coo = mat.tocoo()
    keep = coo.data != 0
    pred = predict(user_f, item_f, coo.row[keep], coo.col[keep], bias).astype(np.float64)
    diff = coo.data[keep].astype(np.float64) - pred
    metric = metric.upper()
    if metric == "MAE":
        return float(np.mean(np.abs(diff)))
    mse = float(np.mean(diff * diff))
    return float(np.sqrt(mse)) if metric == "RMSE" else mse


# --------------------------------------------------------------------------------------
# R2 / R11  WMF.train with early stopping           wmf_model.py:49-189
# --------------------------------------------------------------------------------------
class EarlyStop:
    """Counter of consecutive epochs with mse*(1+min_improvement) > last_mse; last_mse starts
    at -inf so epoch 0 always counts (wmf_model.py:131-132, :164-168, :179-180)."""

    def __init__(self, stopping_rounds, min_improvement):
        self.rounds = stopping_rounds
        self.min_improvement = min_improvement
        self.last = -np.inf
        self.count = 0

    def update(self, mse):
        if mse * (1 + self.min_improvement) > self.last:
            self.count += 1
        else:
            self.count = 0
        self.last = mse
        return self.count >= self.rounds


def train(items0, utility_mat, iterations, eval_mat, count_mat=None, gamma=0.1, weighted=True,
          bias=False, dim=None, alpha=10, stopping_rounds=3, min_improvement=1e-4,
          pre_process_count="log", beta=1, dtype=np.float32):
    """Epoch loop of WMF.train. Returns (users, items, last_epoch_index, mse_trace).
    ``cores`` does not appear: the Pool paths compute the same numbers (R6)."""
    items = np.array(items0, dtype=dtype, copy=True)
    users = None
    trace = []
    stop = EarlyStop(stopping_rounds, min_improvement)
    it = -1
    if weighted is not True:
        if dim is None:
            dim = items.shape[1]
        R = utility_mat.copy()
        for it in range(iterations):
            users = unweighted_half_step(items, R, gamma, dim, dtype)
            items = unweighted_half_step(users, R.T.tocsr(), gamma, dim, dtype)
            mse = eval_prec(users, items, eval_mat, bias)
            trace.append(mse)
            if stop.update(mse):
                break
        return users, items, it, trace
    C = count_mat.copy()
    C.data = preprocess_counts(C.data, pre_process_count, alpha, beta)
    CT = C.T.tocsr()
    step = half_step_bias if bias else half_step
    for it in range(iterations):
        users = step(items, C, gamma, dtype)
        items = step(users, CT, gamma, dtype)
        mse = eval_prec(users, items, eval_mat, bias)
        trace.append(mse)
        if stop.update(mse):
            break
    return users, items, it, trace


# --------------------------------------------------------------------------------------
# R12  RecModel.eval_topn / compute_hit             base_model.py:51-98, :100-148
# --------------------------------------------------------------------------------------
def eval_topn(rank_fn, num_items, test_mat, topn, rand_sampled=1000, random_state=None,
              train_mat=None, eval_mat=None, dtype=np.float32):
    """Sampled Recall@N protocol. ``rank_fn(items, user, topn)`` is the model's rank. Draw order
    from the global NumPy RNG is the reference's: per user with test entries, randint for
    rand_sampled+1 candidate ids (:63) then one slot (:64); each held-out item overwrites the
    slot (:89) before ranking. hits accumulate in ``dtype`` (:125); recall = hits /
    count_nonzero(test_mat) (:143). ``super_mat`` is only used to find users (its += falls
    back to +, so test_mat is not mutated)."""
    if not isinstance(topn, np.ndarray):  # base_model.py:121-122
        raise ValueError("Topn has to be a np.array")
    if random_state is not None:
        np.random.seed(random_state)
    hits = np.zeros(topn.shape, dtype=dtype)
    ptr, ind = test_mat.indptr, test_mat.indices
    kmax = topn.max()
    for user in range(test_mat.shape[0]):
        held = ind[ptr[user]:ptr[user + 1]]
        if len(held) == 0:
            continue
        cand = np.random.randint(0, num_items, size=(rand_sampled + 1))
        slot = np.random.randint(0, rand_sampled - (2 * kmax))
        row_hits = np.zeros(topn.shape, dtype=dtype)
        for item in held:
            cand[slot] = item
            top = rank_fn(cand, user, kmax)
            for p in range(len(topn)):
                if item in top[:topn[p]]:
                    row_hits[p] += 1
        hits += row_hits
    recall = hits / len(test_mat.nonzero()[0])
    return {f"Recall@{topn[p]}": recall[p] for p in range(len(topn))}


# --------------------------------------------------------------------------------------
# §8(d)  algorithmic bytes / flops
# --------------------------------------------------------------------------------------
def half_step_bytes(nnz, rows, cols, f, bias=False):
    """ALGORITHMIC bytes of one half-step:
    nnz*(4f + 4 + 4 [+4 bias gather]) + rows*(4f + 4) + cols*4f."""
    return nnz * (4 * f + 8 + (4 if bias else 0)) + rows * (4 * f + 4) + cols * 4 * f


def epoch_bytes(nnz, users, items, f, bias=False):
    return half_step_bytes(nnz, users, items, f, bias) + half_step_bytes(nnz, items, users, f, bias)


# --------------------------------------------------------------------------------------------
# EASE: restatement of /root/reference/RecModel/ease_model.py:81-114 and
# /root/reference/RecModel/fast_utils/ease_utils.pyx:15-30
# --------------------------------------------------------------------------------------------
def ease_train(X, alpha, dtype=np.float32):
    """W of the EASE model. dtype float32 follows the reference (ease_model.py:93-109); float64 is the yardstick."""
    X_csr = X.copy().tocsr()
    G = np.asarray(np.dot(X_csr.T, X_csr).todense()).astype(dtype)      # :93
    np.fill_diagonal(G, G.diagonal() + dtype(alpha))                    # :97
    res = np.linalg.inv(G)                                              # :102
    res = res / (-res.diagonal() + 1e-9)                                # :106 (column-wise broadcast)
    np.fill_diagonal(res, 0)                                            # :108
    return res.astype(dtype if dtype == np.float64 else np.float32)     # :111


def ease_predict(X, W, users, items):
    """Per pair: sum over the user's stored entries of X[u, j] * W[j, item], FP32 products accumulated in
    double in stored order (ease_utils.pyx:24-29: `output` is a float64 array, the product two C floats)."""
    X_csr = X.tocsr()
    users, items = np.atleast_1d(users), np.atleast_1d(items)
    if len(users) == 0 or len(items) == 0:
        return np.full(1, 0.0, dtype=np.float32)                        # :19-20
    out = np.zeros(len(items))
    W = np.asarray(W, dtype=np.float32)
    for i in range(len(items)):
        u = users[i] if len(users) > 1 else users[0]
        lo, hi = X_csr.indptr[u], X_csr.indptr[u + 1]
        prod = (X_csr.data[lo:hi].astype(np.float32) * W[X_csr.indices[lo:hi], items[i]]).astype(np.float32)
        acc = 0.0
        for p in prod:
            acc += float(p)
        out[i] = acc
    return out


def ease_rank(X, W, items, user, topn):
    """ease_model.py:51-53."""
    items = np.asarray(items)
    pred = ease_predict(X, W, np.full(items.shape[0], user, dtype=np.int32), items.astype(np.int32))
    return items[np.argpartition(pred, list(range(-topn, 0, 1)))[-topn:]][::-1]
