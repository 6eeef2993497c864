"""One resident ALS epoch as four replayable CUDA graphs.

An epoch is a fixed sequence of launches (Gram, schedule table, half-step kernel, conditional fix-up,
and - row-sharded - the exchange of Gram blocks and factor shards). At ML-20M shape on 8 GPUs a half-step
kernel runs for ~1 ms, so ~50 launches and collectives per epoch issued from Python leave the GPU waiting
for the host; captured once, the same sequence replays with four graph launches per epoch.

    user half-step | exchange (Gram of the new user shard, all-gather) | item half-step | exchange

The graphs read and write fixed buffers (``items``, ``users``, ``G_items``, ``G_users``), so epoch k+1's first
graph consumes what epoch k's last graph produced. Events can be recorded between the replays (the
bench times the two half-step kernels that way). wmf_model.py:140-156 is the loop this replaces.
"""
import torch

from . import _lib, engine, sharding


class ResidentEpoch:
    def __init__(self, C, CT, items0, gamma, bias=False, algo=_lib.ALGO_AUTO, ub=None, ib=None, graphs=True):
        """C / CT: this rank's row slices (DeviceCSR) of the count matrix and of its transpose; ub / ib: shard
        boundaries (None on one GPU); items0: full initial item factors on the device."""
        self.C, self.CT, self.gamma, self.bias, self.algo = C, CT, float(gamma), bool(bias), algo
        self.ub, self.ib = ub, ib
        self.world = 1 if ub is None else len(ub) - 1
        dev, f = items0.device, items0.shape[1]
        n_users = C.shape[0] if ub is None else int(ub[-1])
        n_items = CT.shape[0] if ib is None else int(ib[-1])
        self.items = items0.clone()
        self.users = torch.zeros((n_users, f), dtype=torch.float32, device=dev)
        self.G_items = engine.gram(self.items, self.gamma, ones_col0=self.bias)
        self.G_users = torch.zeros_like(self.G_items)
        self.ws = None
        if self.world == 1:   # the new factors ARE the full matrices: the half-steps write them in place
            self.X_users, self.X_items = self.users, self.items
        else:                 # this rank's new shards
            self.X_users = torch.empty((C.shape[0], f), dtype=torch.float32, device=dev)
            self.X_items = torch.empty((CT.shape[0], f), dtype=torch.float32, device=dev)
        C.row_order, CT.row_order, C.split_segments, CT.split_segments  # noqa: B018  (host-side setup, once)
        # The captured graphs hold raw pointers into their scratch: this object owns it (sized once, kept alive
        # with the graphs) instead of borrowing the process-wide grow-only cache of engine.workspace().
        lib = _lib.load()
        need = max(engine.half_step_workspace_bytes(C, f, algo), engine.half_step_workspace_bytes(CT, f, algo),
                   int(lib.wmf_gram_workspace_bytes(max(n_users, n_items), f)))
        self.ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self.stages = [self._user_half_step, self._user_exchange, self._item_half_step, self._item_exchange]
        self.graphs = None
        if graphs:
            self._capture()

    # ---- the four stages (eager form; captured verbatim)
    def _user_half_step(self):
        engine.half_step(self.C, self.items, self.G_items, bias=self.bias, algo=self.algo, out=self.X_users, ws=self.ws)

    def _user_exchange(self):
        if self.world > 1:
            self.G_users.copy_(sharding.sharded_gram(self.X_users, self.ub, self.gamma, ones_col0=self.bias))
            sharding.all_gather_rows(self.X_users, self.ub, out=self.users)
        else:
            engine.gram(self.users, self.gamma, ones_col0=self.bias, ws=self.ws, out=self.G_users)

    def _item_half_step(self):
        engine.half_step(self.CT, self.users, self.G_users, bias=self.bias, algo=self.algo, out=self.X_items, ws=self.ws)

    def _item_exchange(self):
        if self.world > 1:
            self.G_items.copy_(sharding.sharded_gram(self.X_items, self.ib, self.gamma, ones_col0=self.bias))
            sharding.all_gather_rows(self.X_items, self.ib, out=self.items)
        else:
            engine.gram(self.items, self.gamma, ones_col0=self.bias, ws=self.ws, out=self.G_items)

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up on the capture stream: workspaces, attributes, NCCL channels
            for _ in range(2):
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

    # launches of this repo's kernels per epoch (bench.py's gpu_launches): 2 x (tc_maxima, tc_prep_rows,
    # tc_finish_prep, als_half_step_tc, conditional fix-up) + 2 x (gram_partial, gram_reduce)
    KERNELS_PER_EPOCH = 14
