"""One resident ALS epoch as four replayable CUDA graphs.

An epoch is a fixed sequence of launches (Cholesky of the Gram, whitening, schedule tables, the dual and primal
tensor-core kernels, unwhitening, conditional fix-up, Gram of the new factors and - row-sharded - the exchange of
factor rows and Gram blocks). At ML-20M shape on 8 GPUs a half-step runs for well under a millisecond, so the ~25
launches per epoch issued from Python would leave the GPU waiting for the host; captured once, the same sequence
replays with four graph launches per epoch.

    user half-step | exchange (rows + Gram blocks of the new user shard) | item half-step | exchange

Row-sharded, the exchange is the library's own kernel writing into every peer's symmetric copy of the factor and
Gram-partial buffers over NVLink (``sharding.PeerBuffers``) followed by one barrier; where symmetric memory is not
available it falls back to NCCL (all-gather of the rows, all-reduce of zero-padded Gram blocks).

The graphs read and write fixed buffers (``items``, ``users``, ``G_items``, ``G_users``), so epoch k+1's first
graph consumes what epoch k's last graph produced. Events can be recorded between the replays (the bench times the
two half-step stages that way). wmf_model.py:140-156 is the loop this replaces.
"""
import torch

from . import _lib, engine, sharding


class ResidentEpoch:
    def __init__(self, C, CT, items0, gamma, bias=False, algo=_lib.ALGO_AUTO, ub=None, ib=None, graphs=True, peer=True,
                 count_launches=True, shared_ws=False):
        """C / CT: this rank's row slices (DeviceCSR) of the count matrix and of its transpose; ub / ib: shard
        boundaries (None on one GPU); items0: full initial item factors on the device; peer: exchange over peer
        memory when it is available (else NCCL)."""
        self.C, self.CT, self.gamma, self.bias, self.algo = C, CT, float(gamma), bool(bias), algo
        self.ub, self.ib = ub, ib
        self.world = 1 if ub is None else len(ub) - 1
        self.rank = sharding.dist_info()[0] if self.world > 1 else 0
        dev, f = items0.device, items0.shape[1]
        n_users = C.shape[0] if ub is None else int(ub[-1])
        n_items = CT.shape[0] if ib is None else int(ib[-1])
        self.n_users, self.n_items = n_users, n_items
        self.px = sharding.PeerBuffers.cached(n_users, n_items, f, dev) if (self.world > 1 and peer) else None
        if self.px is not None:
            self.items, self.users = self.px.views["items"], self.px.views["users"]
            self.items.copy_(items0)
            self.users.zero_()
            self.exchange_mode = "peer-memory stores (wmf_peer_broadcast) + 1 barrier per half-step"
        else:
            self.items = items0.clone()
            self.users = torch.zeros((n_users, f), dtype=torch.float32, device=dev)
            self.exchange_mode = "none (single GPU)" if self.world == 1 else "nccl all-gather + all-reduce"
        self.G_items = engine.gram(self.items, self.gamma, ones_col0=self.bias)
        self.G_users = torch.zeros_like(self.G_items)
        if self.world == 1:   # the new factors ARE the full matrices: the half-steps write them in place
            self.X_users, self.X_items = self.users, self.items
        elif self.px is not None:   # this rank's rows of the symmetric full matrices
            self.X_users = self.users[int(ub[self.rank]):int(ub[self.rank + 1])]
            self.X_items = self.items[int(ib[self.rank]):int(ib[self.rank + 1])]
        else:                 # this rank's new shards
            self.X_users = torch.empty((C.shape[0], f), dtype=torch.float32, device=dev)
            self.X_items = torch.empty((CT.shape[0], f), dtype=torch.float32, device=dev)
        C.row_order, CT.row_order, C.split_segments, CT.split_segments  # noqa: B018  (host-side setup, once)
        # The captured graphs hold raw pointers into their scratch: this object owns it (sized once, kept alive
        # with the graphs) instead of borrowing the process-wide grow-only cache of engine.workspace().
        lib = _lib.load()
        need = max(engine.half_step_workspace_bytes(C, f, algo), engine.half_step_workspace_bytes(CT, f, algo),
                   int(lib.wmf_gram_workspace_bytes(max(n_users, n_items), f)))
        # (an eager, un-captured epoch - WMF.train - may borrow the cache instead: a fresh ~0.5 GB allocation per
        # train() call intermittently misses the caching allocator and costs 10 ms of cudaMalloc)
        if shared_ws and not graphs:
            self.ws = engine.workspace(need, dev)
        else:
            self.ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._side = torch.cuda.Stream(device=dev) if self.px is not None else None
        self._fork, self._join = torch.cuda.Event(), torch.cuda.Event()
        self.stages = [self._user_half_step, self._user_exchange, self._item_half_step, self._item_exchange]
        self.graphs = None
        self.launches_per_epoch = 0
        if self.px is not None:
            # every rank's initial items / zeroed users are in place (and nobody still reads the buffers of an earlier
            # epoch object) before anyone pushes
            torch.cuda.synchronize(dev)
            self.px.barrier()
        if graphs:
            self._capture()
        elif count_launches:
            self._count_launches()

    # ---- the four stages (eager form; captured verbatim)
    def _user_half_step(self):
        engine.half_step(self.C, self.items, self.G_items, bias=self.bias, algo=self.algo, out=self.X_users, ws=self.ws)

    def _exchange(self, X, bounds, n_total, name, G_out, full):
        if self.px is not None:
            lo, hi = int(bounds[self.rank]), int(bounds[self.rank + 1])
            B = engine.gram_block_rows(n_total)
            gp = self.px.views["gp_" + name]
            # the rows travel over NVLink on a side stream while the Gram blocks of the shard are computed
            cur = torch.cuda.current_stream(X.device)
            self._fork.record(cur)
            self._side.wait_event(self._fork)
            with torch.cuda.stream(self._side):
                self.px.push(name, lo, hi)                                    # new factor rows -> every peer
                self._join.record(self._side)
            engine.gram_partials(X, lo, n_total, ones_col0=self.bias, out=gp)  # Gram blocks of the shard
            self.px.push("gp_" + name, lo // B, -(-hi // B))                   # ... -> every peer
            cur.wait_event(self._join)
            self.px.barrier()
            engine.gram_from_partials(gp, n_total, self.gamma, out=G_out)      # all blocks in block order
        elif self.world > 1:
            G_out.copy_(sharding.sharded_gram(X, bounds, self.gamma, ones_col0=self.bias))
            sharding.all_gather_rows(X, bounds, out=full)
        else:
            engine.gram(full, self.gamma, ones_col0=self.bias, ws=self.ws, out=G_out)

    def _user_exchange(self):
        self._exchange(self.X_users, self.ub, self.n_users, "users", self.G_users, self.users)

    def _item_half_step(self):
        engine.half_step(self.CT, self.users, self.G_users, bias=self.bias, algo=self.algo, out=self.X_items, ws=self.ws)

    def _item_exchange(self):
        self._exchange(self.X_items, self.ib, self.n_items, "items", self.G_items, self.items)

    def _count_launches(self):
        lib = _lib.load()
        n0 = int(lib.wmf_launch_count())
        for st in self.stages:
            st()
        self.launches_per_epoch = int(lib.wmf_launch_count()) - n0

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up on the capture stream: workspaces, attributes, NCCL channels
            self._count_launches()
            for st in self.stages:
                st()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs = []
        pool = None
        for st in self.stages:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                st()
            pool = g.pool()
            self.graphs.append(g)

    def run_stage(self, k):
        if self.graphs is not None:
            self.graphs[k].replay()
        else:
            self.stages[k]()

    def step(self, events=None):
        """One epoch. ``events``: optional list of 4 CUDA events recorded around the two half-step stages."""
        if events is not None:
            events[0].record()
        self.run_stage(0)
        if events is not None:
            events[1].record()
        self.run_stage(1)
        if events is not None:
            events[2].record()
        self.run_stage(2)
        if events is not None:
            events[3].record()
        self.run_stage(3)
