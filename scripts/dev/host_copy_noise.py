"""Distribution of the staging copy time on the GPU box's host (development): torch's parallel copy at several thread
counts against a single-threaded NumPy copy, 40 repetitions each with the pauses an e2e step has between uploads."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
n = 16_000_000
a = np.random.default_rng(0).integers(0, 1000, n).astype(np.int32)
buf = torch.empty(n, dtype=torch.int32, pin_memory=True)
t = torch.from_numpy(a)
def dist(label, fn, reps=40, pause=0.015):
    ts = []
    for _ in range(reps):
        time.sleep(pause)
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    ts = np.array(ts)
    print(f"{label:34s} min {ts.min():6.2f}  median {np.median(ts):6.2f}  p90 {np.percentile(ts, 90):6.2f}  max {ts.max():6.2f} ms", flush=True)
for k in (16, 8, 4, 2, 1):
    torch.set_num_threads(k)
    dist(f"torch copy_ {k} threads", lambda: buf.copy_(t))
dist("np.copyto single", lambda: np.copyto(buf.numpy(), a))
