import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch
from oracle import wmf_oracle as orc
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR
from recmodel_b200.synthetic import make_counts_cached
C = make_counts_cached(6040,3706,1_000_000, seed=31); C.data = orc.preprocess_counts(C.data)
f=64; Y = orc.init_items(3706, f, False)
rows=slice(0,1500)
x32 = orc.half_step(Y, C[rows], 0.1, np.float32); x64 = orc.half_step(Y, C[rows], 0.1, np.float64)
dev=torch.device('cuda:0')
Yd=torch.from_numpy(Y).to(dev); Cd=DeviceCSR.from_scipy(C,dev)
G=engine.gram(Yd,0.1)
X=engine.half_step(Cd,Yd,G,algo=_lib.ALGO_SIMT).cpu().numpy()[rows]
def rel(a,b): return np.linalg.norm(a-b,axis=1)/np.linalg.norm(b,axis=1)
eg=rel(X,x64); er=rel(x32,x64)
nnz=np.diff(C.indptr)[rows]
order=np.argsort(-eg)[:12]
for r in order: print(r, nnz[r], 'gpu %.2e ref %.2e'%(eg[r],er[r]), 'norm', np.linalg.norm(x64[r]))
print('median gpu %.2e ref %.2e'%(np.median(eg),np.median(er)))
G64=Y.astype(np.float64).T@Y.astype(np.float64)+0.1*np.eye(f)
print('G err', np.abs(G.cpu().numpy()-G64).max()/G64.max())
# corr of error with nnz
for lo,hi in ((1,20),(20,60),(60,150),(150,400),(400,5000)):
    m=(nnz>=lo)&(nnz<hi)
    if m.any(): print(lo,hi,m.sum(),'gpu max %.2e med %.2e | ref max %.2e med %.2e'%(eg[m].max(),np.median(eg[m]),er[m].max(),np.median(er[m])))
