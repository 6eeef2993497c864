import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from recmodel_b200 import engine
n = 16_000_000
a = np.random.default_rng(0).integers(0, 1000, n).astype(np.int32)
buf = torch.empty(n, dtype=torch.int32, pin_memory=True)
t = torch.from_numpy(a)
def T(label, fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(f"rank {os.environ.get('RANK', 0)} threads(torch)={torch.get_num_threads()} OMP={os.environ.get('OMP_NUM_THREADS')} cpus={os.cpu_count()} {label:40s} min {min(ts)*1e3:6.2f} ms  median {sorted(ts)[len(ts)//2]*1e3:6.2f} ms", flush=True)
T("torch copy_ (current threads)", lambda: buf.copy_(t))
T("np.copyto single", lambda: np.copyto(buf.numpy(), a))
T("engine._host_copy", lambda: engine._host_copy(buf, t))
old = torch.get_num_threads()
for k in (4, 8, 16):
    torch.set_num_threads(k)
    T(f"torch copy_ set_num_threads({k})", lambda: buf.copy_(t))
torch.set_num_threads(old)
