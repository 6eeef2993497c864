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


class PeerBuffers:
    """Symmetric (peer-mapped) copies of everything a row-sharded epoch exchanges: the full user and item factor
    matrices and the Gram block partials of both sides. Rank g writes the rows / blocks it owns into every peer's
    copy with the library's own kernel (``engine.peer_broadcast`` -> wmf_peer_broadcast: posted stores over
    NVLink / NVSwitch); one barrier per half-step orders the stores before the readers. No NCCL all-gather, no
    all-reduce, and every rank adds ALL Gram blocks in block order: the bits of the single-GPU Gram.
    ``create`` returns None where symmetric memory is unavailable (CPU / gloo tests): callers fall back to
    ``all_gather_rows`` / ``sharded_gram``."""

    def __init__(self, handle, flat, views, table, rank, world):
        self.handle, self.flat, self.views, self.table, self.rank, self.world = handle, flat, views, table, rank, world

    _cache = {}

    @classmethod
    def cached(cls, n_users, n_items, f, device, group=None):
        """``create`` once per (shape, device): the rendezvous (IPC handle exchange) costs milliseconds and the
        buffers can be reused by consecutive trainings of the same shape."""
        key = (int(n_users), int(n_items), int(f), str(device), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls.create(n_users, n_items, f, device, group)
        return cls._cache[key]

    @classmethod
    def create(cls, n_users, n_items, f, device, group=None):
        from . import engine
        rank, world = dist_info(group)
        if world == 1 or device.type != "cuda":
            return None
        try:
            import torch.distributed._symmetric_memory as symm
        except Exception:
            return None
        bu, bi = engine.gram_blocks(n_users), engine.gram_blocks(n_items)
        sizes = {"users": n_users * f * 4, "items": n_items * f * 4, "gp_users": bu * f * f * 8, "gp_items": bi * f * f * 8}
        offs, total = {}, 0
        for k, nb in sizes.items():
            offs[k] = total
            total += (nb + 255) // 256 * 256
        try:
            flat = symm.empty(total, dtype=torch.uint8, device=device)
            grp = group if group is not None else dist.group.WORLD
            try:
                handle = symm.rendezvous(flat, group=grp)
            except TypeError:
                handle = symm.rendezvous(flat, grp.group_name)
            ptrs = [int(x) for x in handle.buffer_ptrs]
        except Exception as exc:  # no peer access / unsupported allocator: NCCL path
            import warnings
            warnings.warn(f"symmetric memory unavailable ({type(exc).__name__}: {exc}); exchanging factor shards with NCCL")
            return None
        views = {
            "users": flat[offs["users"]:offs["users"] + sizes["users"]].view(torch.float32).view(n_users, f),
            "items": flat[offs["items"]:offs["items"] + sizes["items"]].view(torch.float32).view(n_items, f),
            "gp_users": flat[offs["gp_users"]:offs["gp_users"] + sizes["gp_users"]].view(torch.float64).view(bu, f, f),
            "gp_items": flat[offs["gp_items"]:offs["gp_items"] + sizes["gp_items"]].view(torch.float64).view(bi, f, f),
        }
        table = torch.tensor(ptrs, dtype=torch.int64, device=device)
        out = cls(handle, flat, views, table, rank, world)
        out.offs = offs
        return out

    def push(self, name, lo, hi):
        """Rows (factors) or blocks (Gram partials) [lo, hi) of this rank's copy of ``name`` -> every peer's copy."""
        from . import engine
        v = self.views[name]
        if hi <= lo:
            return
        part = v[lo:hi]
        row_bytes = part[0].numel() * part.element_size()
        engine.peer_broadcast(part, self.table, self.world, self.rank, self.offs[name] + lo * row_bytes)

    def barrier(self):
        self.handle.barrier(channel=0)
