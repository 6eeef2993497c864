"""BASELINE.json config 3 on one GPU: WMF dim 128 at Netflix shape (480 189 x 17 770, 100 M stored entries).
The count matrix is generated on the device (same recipe as recmodel_b200.synthetic.make_counts: log-normal user
activity, item popularity ~ rank^-0.8, distinct pairs, counts 1..5) because the host generator needs minutes at
this size. Reports the epoch time of the resident loop and checks sampled rows against an fp64 solve."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR
from recmodel_b200.epoch import ResidentEpoch

U, I, NNZ, F = 480_189, 17_770, 100_000_000, 128
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(20240229)
act = torch.exp(torch.randn(U, device=dev, generator=g, dtype=torch.float64)); cum_u = torch.cumsum(act / act.sum(), 0)
pop = torch.arange(1, I + 1, device=dev, dtype=torch.float64) ** -0.8
pop = pop[torch.randperm(I, device=dev, generator=g)]; cum_i = torch.cumsum(pop / pop.sum(), 0)
keys = torch.empty(0, dtype=torch.int64, device=dev)
while keys.numel() < NNZ:
    m = int((NNZ - keys.numel()) * 1.3) + 1024
    u = torch.searchsorted(cum_u, torch.rand(m, device=dev, generator=g, dtype=torch.float64)).clamp_(max=U - 1)
    i = torch.searchsorted(cum_i, torch.rand(m, device=dev, generator=g, dtype=torch.float64)).clamp_(max=I - 1)
    keys = torch.unique(torch.cat([keys, u * I + i]))
keys = keys[torch.randperm(keys.numel(), device=dev, generator=g)[:NNZ]].sort().values
rows, cols = keys // I, (keys % I).to(torch.int32)
indptr = torch.zeros(U + 1, dtype=torch.int64, device=dev); indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=U), 0)
data = torch.randint(1, 6, (NNZ,), device=dev, generator=g).to(torch.float32)
del keys, rows, u, i
C = DeviceCSR(indptr, cols, data, (U, I))
engine.preprocess_(C.data, "log", 10, 1)
CT = C.transpose()
cu, ci = (C.indptr[1:] - C.indptr[:-1]), (CT.indptr[1:] - CT.indptr[:-1])
print(f"matrix {U} x {I}, {C.nnz} entries; longest user row {int(cu.max())}, longest item row {int(ci.max())}; "
      f"segments users {C.split_segments} items {CT.split_segments}", flush=True)
items0 = torch.rand((I, F), device=dev, generator=g)
loop = ResidentEpoch(C, CT, items0, 0.1, graphs=True)
for _ in range(2):
    loop.step()
torch.cuda.synchronize()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(3)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(3):
    loop.step(ev[k])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
tu = np.mean([e[0].elapsed_time(e[1]) for e in ev]); ti = np.mean([e[2].elapsed_time(e[3]) for e in ev])
bytes_epoch = 2 * C.nnz * (4 * F + 8) + (U + I) * (4 * F + 4)
print(f"epoch {ms:.2f} ms = {2 * C.nnz / ms / 1e6:.2f} G nnz-updates/s (user half-step {tu:.2f} ms, item half-step {ti:.2f} ms); "
      f"algorithmic {bytes_epoch / 1e9:.1f} GB -> {bytes_epoch / ms / 1e6:.0f} GB/s", flush=True)
# accuracy of sampled rows of one more item half-step against fp64 (heaviest rows included)
G = engine.gram(loop.users, 0.1)
X = engine.half_step(CT, loop.users, G).cpu().numpy()
assert np.all(np.isfinite(X))
Uh = loop.users.cpu().numpy().astype(np.float64)
G64 = Uh.T @ Uh + 0.1 * np.eye(F)
order = torch.argsort(ci, descending=True).cpu().numpy()
ip, idx, dat = CT.indptr.cpu().numpy(), CT.indices.cpu().numpy(), CT.data.cpu().numpy()
worst = 0.0
for r in list(order[:3]) + list(order[1000:1002]) + list(order[10000:10002]):
    lo, hi = ip[r], ip[r + 1]
    Yr = Uh[idx[lo:hi]]; d = dat[lo:hi].astype(np.float64)
    x = np.linalg.solve(G64 + (Yr * d[:, None]).T @ Yr, (d + 1) @ Yr)
    err = np.linalg.norm(X[r] - x) / np.linalg.norm(x)
    worst = max(worst, err)
    print(f"  item row {r}: {hi - lo} entries, error vs fp64 {err:.2e}")
print(f"worst {worst:.2e}")
assert worst < 1e-4
print("netflix shape ok")
