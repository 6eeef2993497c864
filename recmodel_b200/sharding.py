"""Row partitioning and factor exchange for multi-GPU ALS.

Within a half-step every row's solve is independent given the full fixed-side factors
(wmf_model.py:220-239), so rank g owns a contiguous row range of the count matrix (user
half-step) and of its transpose (item half-step); the only exchange is an all-gather of the
new factor shard after each half-step (NCCL over NVLink/NVSwitch on GPUs, gloo on CPU in the
tests). Row -> rank assignment never changes a row's arithmetic, so N-GPU factors equal
1-GPU factors bit for bit.
"""
import numpy as np
import torch
import torch.distributed as dist


def dist_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def row_costs(counts, f):
    """Per-row work model: n_r * f^2 (Gram) + f^3/3 (solve) (SURVEY.md §8d flop terms)."""
    counts = np.asarray(counts, dtype=np.float64)
    return counts * float(f) * f + np.where(counts > 0, float(f) ** 3 / 3.0, 1.0)


def balanced_row_partition(counts, world, f, align=1):
    """Contiguous row ranges [b[g], b[g+1]) with near-equal summed cost. Returns int64 array of
    world+1 boundaries (monotone, b[0]=0, b[-1]=rows). Inner boundaries are rounded to multiples of
    ``align`` (the Gram block height, so that every rank owns whole Gram blocks)."""
    counts = np.asarray(counts)
    rows = len(counts)
    if world <= 1 or rows == 0:
        return np.array([0] + [rows] * max(world, 1), dtype=np.int64)
    csum = np.cumsum(row_costs(counts, f))
    targets = csum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(csum, targets, side="left") + 1
    if align > 1:
        cuts = (cuts + align // 2) // align * align
    bounds = np.concatenate([[0], np.minimum(cuts, rows), [rows]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def sharded_gram(X_local, bounds, lam, ones_col0=False, group=None):
    """G = X^T X + lam I of the full factor matrix from this rank's freshly computed shard: block partials
    of the local rows, one sum all-reduce (foreign blocks are zeros, so the sum is exact), blocks added in
    block order. Same bits as the single-GPU Gram of the gathered matrix."""
    from . import engine
    rank, world = dist_info(group)
    n_total = int(bounds[-1])
    partials = engine.gram_partials(X_local, int(bounds[rank]), n_total, ones_col0=ones_col0)
    if world > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return engine.gram_from_partials(partials, n_total, lam)


def all_gather_rows(local, bounds, group=None, out=None):
    """Concatenate row shards [bounds[g], bounds[g+1]) from every rank into the full matrix. Shards have
    different heights: the gather writes every shard straight into its rows of ``out`` (allocated if None), no
    padding and no repacking copies."""
    rank, world = dist_info(group)
    if world == 1:
        if out is None:
            return local
        out.copy_(local)
        return out
    f = local.shape[1]
    if out is None:
        out = torch.empty((int(bounds[-1]), f), dtype=local.dtype, device=local.device)
    heights = np.diff(bounds)
    if dist.get_backend(group) == "nccl" or int(heights.min()) == int(heights.max()):
        views = [out[int(bounds[g]):int(bounds[g + 1])] for g in range(world)]
        dist.all_gather(views, local.contiguous(), group=group)
        return out
    # backends without uneven all-gather (gloo, in the CPU tests): pad to the tallest shard, gather, unpack
    hmax = int(heights.max())
    send = torch.zeros((hmax, f), dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    recv = torch.empty((world * hmax, f), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    for g in range(world):
        out[int(bounds[g]):int(bounds[g + 1])] = recv[g * hmax: g * hmax + int(heights[g])]
    return out


def all_reduce_sum_(t, group=None):
    _, world = dist_info(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
