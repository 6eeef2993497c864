#!/usr/bin/env python
"""bench.py - WMF nnz-updates/s per epoch at ML-20M shape (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU arithmetic

A "step" is one ALS epoch (user half-step + item half-step, each with its Gram) over the
synthetic 138 493 x 26 744 count matrix with 20 M stored entries, dim 128, weighted, log
preprocessing; metric = 2*nnz / t_epoch. `value` is measured with the matrices resident in
HBM; `e2e` is the same metric through the public `WMF.train` call with HOST (SciPy/NumPy)
buffers, so it pays the H2D upload, the device transpose, the epoch, the fused eval and the
D2H read-back of the factors every step. Under torchrun the rows are partitioned across the
ranks (strong scaling: same matrix at every N) and factor shards are all-gathered after every
half-step over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: users, items, nnz, dim
    "ml20m": (138_493, 26_744, 20_000_000, 128),
    "ml1m": (6040, 3706, 1_000_000, 64),      # debugging only; never the reported config
}
GAMMA, ALPHA, BETA = 0.1, 10, 1


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_traffic():
    """DRAM bytes (read + write) of the two half-step launches of one epoch from the committed
    `ncu --set full` capture (profiles/r01c_ncu_traffic.json); None if absent."""
    path = os.path.join(ROOT, "profiles", "r01c_ncu_traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh)
        return float(t["user_half_step_dram_bytes"]) + float(t["item_half_step_dram_bytes"])
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                p = json.load(fh)
            return float(p["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth(workload, planted=0):
    from recmodel_b200.synthetic import make_counts_cached
    users, items, nnz, dim = WORKLOADS[workload]
    return make_counts_cached(users, items, nnz, planted_rank=planted), dim


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's arithmetic (oracle port, NumPy + BLAS + LAPACK, all host threads)
# --------------------------------------------------------------------------------------------
def cpu_epoch_sample(C, CT, dim, frac_u, frac_i, threads):
    """Time the oracle's two half-steps on contiguous row blocks holding ~frac of the rows;
    extrapolate linearly in rows to one epoch. Returns (nnz-updates/s, description)."""
    from oracle import wmf_oracle as orc
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None
    users, items = C.shape
    nu = max(1, int(users * frac_u))
    ni = max(1, int(items * frac_i))
    Y = orc.init_items(items, dim, False)
    ctx = threadpool_limits(limits=threads) if threadpool_limits else None
    try:
        t0 = time.perf_counter()
        Xu = orc.half_step(Y, C[:nu], GAMMA)
        t_u = time.perf_counter() - t0
        # item side needs user factors for all users: reuse the sample's rows cyclically (same cost)
        Ufull = np.resize(Xu, (users, dim)).astype(np.float32, copy=False)
        t0 = time.perf_counter()
        orc.half_step(Ufull, CT[:ni], GAMMA)
        t_i = time.perf_counter() - t0
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None)
    t_epoch = t_u * users / nu + t_i * items / ni
    desc = (f"oracle port (NumPy/BLAS/LAPACK per-row loop of wmf_model.py:213-240): {nu} of {users} user rows in "
            f"{t_u:.2f}s + {ni} of {items} item rows in {t_i:.2f}s, extrapolated linearly in rows to one epoch")
    return 2.0 * C.nnz / t_epoch, t_epoch, desc


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return  # under torchrun only rank 0 measures the CPU arm
    from oracle import wmf_oracle as orc
    C, dim = synth(args.workload)
    C = C.copy()
    C.data = orc.preprocess_counts(C.data, "log", ALPHA, BETA)
    CT = C.T.tocsr()
    threads = os.cpu_count() or 1
    # bounded sample per step: ~10 % of the rows each way (a few seconds of CPU work per step)
    frac = args.cpu_frac if args.cpu_frac else 0.1
    for _ in range(args.warmup):
        cpu_epoch_sample(C, CT, dim, frac / 4, frac / 4, threads)
    vals, t_eps, desc = [], [], ""
    for _ in range(args.steps):
        v, t_ep, desc = cpu_epoch_sample(C, CT, dim, frac, frac, threads)
        vals.append(v)
        t_eps.append(t_ep)
    users, items, nnz, _ = WORKLOADS[args.workload]
    value = 2.0 * C.nnz / float(np.mean(t_eps))  # consistent with ms_per_step
    line = {
        "impl": "reference", "metric": "wmf_nnz_updates_per_sec_per_epoch", "value": value, "unit": "nnz-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(t_eps)) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"WMF weighted ALS epoch, ML-20M shape {users}x{items}, {nnz} nnz, dim {dim}, "
                               "log preprocessing (BASELINE.json configs[1])" if args.workload == "ml20m"
                   else args.workload, "l2": "n/a (cpu)"},
        "cpu_baseline": {"value": value, "unit": "nnz-updates/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "nnz-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from recmodel_b200 import WMF, _lib, engine, sharding
    from recmodel_b200.engine import DeviceCSR
    from recmodel_b200.synthetic import split_train_test
    from oracle import wmf_oracle as orc  # cpu_baseline leg + algorithmic byte counts only

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    _lib.require_device()
    algo = {"auto": _lib.ALGO_AUTO, "simt": _lib.ALGO_SIMT, "tcgen05": _lib.ALGO_TCGEN05}[args.algo]

    # synthetic data: rank 0 generates (cached on disk), the others wait and load the cache
    if rank == 0:
        C_host, dim = synth(args.workload)
    if world > 1:
        dist.barrier()
    if rank != 0:
        C_host, dim = synth(args.workload)
    users, items = C_host.shape
    nnz = C_host.nnz
    f = dim

    # ---- resident setup (not timed): upload, preprocess, transpose, shard
    C_full = DeviceCSR.from_scipy(C_host, device)
    engine.preprocess_(C_full.data, "log", ALPHA, BETA)
    CT_full = C_full.transpose()
    if world > 1:
        ub = sharding.balanced_row_partition(np.diff(C_host.indptr), world, f, align=engine.gram_block_rows(users))
        ib = sharding.balanced_row_partition((CT_full.indptr[1:] - CT_full.indptr[:-1]).cpu().numpy(), world, f,
                                             align=engine.gram_block_rows(items))
        C = C_full.row_slice(int(ub[rank]), int(ub[rank + 1]))
        CT = CT_full.row_slice(int(ib[rank]), int(ib[rank + 1]))
    else:
        ub = ib = None
        C, CT = C_full, CT_full
    C.row_order, CT.row_order  # noqa: B018
    del C_full, CT_full
    items_d = torch.from_numpy(orc.init_items(items, dim, False)).to(device)
    from recmodel_b200.epoch import ResidentEpoch
    # the epoch as four replayable CUDA graphs (half-step | exchange | half-step | exchange); --no-graphs
    # launches the same sequence from Python
    launch_mode = "python" if args.no_graphs else "4 CUDA graphs per epoch"
    try:
        loop = ResidentEpoch(C, CT, items_d, GAMMA, bias=False, algo=algo, ub=ub, ib=ib, graphs=not args.no_graphs)
    except Exception as exc:  # graph capture refused (driver / NCCL combination): same launches issued from Python
        if args.no_graphs:
            raise
        print(f"[bench] CUDA graph capture failed ({type(exc).__name__}: {exc}); launching the epoch from Python",
              file=sys.stderr, flush=True)
        torch.cuda.synchronize(device)
        launch_mode = "python (graph capture failed)"
        loop = ResidentEpoch(C, CT, items_d, GAMMA, bias=False, algo=algo, ub=ub, ib=ib, graphs=False)
    ev_pairs = []

    def epoch(record=False):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        loop.step(e)
        if record:
            ev_pairs.append(e)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        epoch()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        epoch(record=True)
    stop.record()
    sync_all()
    elapsed_ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = float(elapsed_ms.item()) / args.steps
    value = 2.0 * nnz / (ms_per_step * 1e-3)

    # dominant kernel = the half-step kernel (two launches per epoch); device time per launch
    t_user = float(np.mean([e[0].elapsed_time(e[1]) for e in ev_pairs]))
    t_item = float(np.mean([e[2].elapsed_time(e[3]) for e in ev_pairs]))
    # algorithmic bytes of THIS rank's two launches (without the Gram's read of Y,
    # which belongs to the separate Gram kernel)
    bytes_user = C.nnz * (4 * f + 8) + C.shape[0] * (4 * f + 4)
    bytes_item = CT.nnz * (4 * f + 8) + CT.shape[0] * (4 * f + 4)
    peak, peak_kind = load_peaks()
    achieved = (bytes_user + bytes_item) / ((t_user + t_item) * 1e-3) / 1e9
    flops = (C.nnz + CT.nnz) * (2.0 * f * f + 2 * f) + (C.shape[0] + CT.shape[0]) * (f ** 3 / 3.0 + 2.0 * f * f)

    # ---- e2e through the public API with host buffers (rank-local timing, max over ranks)
    tr_host, te_host = split_train_test(C_host, train=0.8, seed=1993)
    model = WMF(num_items=items, num_users=users, dim=dim, gamma=GAMMA, weighted=True, bias=False, device=device,
                algo=args.algo)
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        t0 = time.perf_counter()
        model.train(tr_host, iterations=1, eval_mat=te_host, count_mat=tr_host, cores=1, stopping_rounds=99)
        u, i = model.users, model.items  # D2H of the result
        return time.perf_counter() - t0, u.nbytes + i.nbytes + 24

    times, d2h = [], 0
    if not args.no_e2e:
        e2e_step()  # warm-up (allocator, pinned staging)
        sync_all()
        e2e_step()  # second warm-up: pinned staging / host allocator caches reach steady state
        sync_all()
        for _ in range(e2e_steps):
            sync_all()
            t, d2h = e2e_step()
            times.append(t)
    else:
        times = [float("nan")]
    t_e2e = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = 2.0 * tr_host.nnz / float(t_e2e.item())
    h2d = (tr_host.nnz * 8 + (users + 1) * 8) + (te_host.nnz * 8 + (users + 1) * 8)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            Cp = C_host.copy()
            Cp.data = orc.preprocess_counts(Cp.data, "log", ALPHA, BETA)
            threads = os.cpu_count() or 1
            v, _, desc = cpu_epoch_sample(Cp, Cp.T.tocsr(), dim, 0.3, 0.3, threads)  # ~10-15 s of host work
            cpu = {"value": v, "unit": "nnz-updates/s", "cores": threads, "kind": "port", "sample": desc}
        line = {
            "metric": "wmf_nnz_updates_per_sec_per_epoch", "value": value, "unit": "nnz-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"WMF weighted ALS epoch, ML-20M shape {users}x{items}, {nnz} nnz, dim {dim}, "
                                    "log preprocessing (BASELINE.json configs[1])") if args.workload == "ml20m"
                       else args.workload,
                       "l2": "inputs larger than L2 (2 x 160 MB CSR + 85 MB factors streamed per epoch)",
                       "algo": args.algo, "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
                       "launch": launch_mode},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic() if world == 1 else None,
                         "traffic_note": "DRAM read+write bytes of the same two launches (ncu --set full, profiles/); "
                                         "algorithmic bytes = %d" % (bytes_user + bytes_item),
                         "peak_kind": peak_kind, "kernel": "als_half_step (2 launches/epoch)",
                         "ms_user_half_step": t_user, "ms_item_half_step": t_item,
                         "fp32_equiv_tflops": flops / ((t_user + t_item) * 1e-3) / 1e12},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "nnz-updates/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": float(t_e2e.item()) * 1e3,
                    "ms_steps_rank0": [round(t * 1e3, 2) for t in times],
                    "what": "WMF.train(host CSR, iterations=1) incl. upload, preprocess, transpose, epoch, eval_prec, "
                            "factor read-back; 80/20 split so nnz = train nnz"},
            # per epoch: 2 x (gram_partial, gram_reduce, tc_maxima, tc_prep_rows, tc_finish_prep, als_half_step_tc,
            # conditional SIMT fix-up)
            "gpu_launches": 14 * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured NCCL collectives keep the communicator busy at teardown (destroy_process_group was seen to
        # hang with live graphs): drop the graphs, drain the device, meet once more and leave without it.
        loop.graphs = None
        del loop
        torch.cuda.synchronize(device)
        dist.barrier()
        torch.cuda.synchronize(device)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="ml20m")
    ap.add_argument("--algo", choices=["auto", "simt", "tcgen05"], default="auto")
    ap.add_argument("--cpu-frac", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel tuning runs: skip the end-to-end leg (the line's e2e is NaN)")
    ap.add_argument("--no-graphs", action="store_true", help="launch the epoch from Python instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
