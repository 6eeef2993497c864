"""Per-role cycle breakdown of CTA 0 of the tcgen05 half-step kernel (WMF_TC_PROFILE=1)."""
import os, sys
os.environ["WMF_TC_PROFILE"] = "1"
sys.path.insert(0, ".")
import numpy as np, torch
from oracle import wmf_oracle as orc
from recmodel_b200 import engine, _lib
from recmodel_b200.engine import DeviceCSR
from recmodel_b200.synthetic import make_counts_cached
C = make_counts_cached(138493, 26744, 20_000_000)
dev = torch.device("cuda:0")
Cd = DeviceCSR.from_scipy(C, dev); engine.preprocess_(Cd.data, "log", 10, 1); CT = Cd.transpose()
Y = torch.from_numpy(orc.init_items(26744, 128, False)).to(dev)
def run(csr, Yd, name):
    G = engine.gram(Yd, 0.1)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); X = engine.half_step(csr, Yd, G, algo=_lib.ALGO_TCGEN05); e1.record(); torch.cuda.synchronize()
    ws = engine.workspace(0, dev)
    prof = ws[256:256 + 640].view(torch.int64).cpu().numpy()
    flags = ws[0:8].view(torch.int32).cpu().numpy()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {ms:.2f} ms  flags={flags}")
    f = lambda c: f"{c/1.965e6:.2f}ms"
    print("  gather total", f(prof[0]), "wait_empty", f(prof[1]), "wait_bempty", f(prof[2]), "chunks", prof[3], "rows", prof[4])
    print("  mma total", f(prof[8]), "wait_full", f(prof[9]), "wait_accempty", f(prof[10]))
    print("  solver0 total", f(prof[16]), "wait(b,acc)full", f(prof[17]), "factor", f(prof[18]), "backsub", f(prof[19]), "rows", prof[22])
    print("  gather phases: issue", f(prof[11]), "wait_stg", f(prof[5]), "xform", f(prof[12]), "arrive", f(prof[13]))
    print("  solver phases (thread 0 of group 0): wait_mma", f(prof[24]), "tmem_ld", f(prof[25]), "own_factor(4 of 16 panels)", f(prof[26]),
          "barA", f(prof[27]), "P", f(prof[28]), "split+sts+fence", f(prof[29]), "barB", f(prof[30]), "mma_issue", f(prof[31]))
    print("  cg: rows solved", prof[32], "of", prof[22], " products", prof[33])
    print("  cg phases (thread 0, group 0), primal: publish+barrier", f(prof[48]), "product", f(prof[49]), "reduce", f(prof[50]),
          "| dual: publish+barrier", f(prof[52]), "product", f(prof[53]), "reduce", f(prof[54]))
    if prof[44] > 0:
        print("  dual kernel, group 0 of CTA 0: total", f(prof[40]), "wait acc_full", f(prof[41]), "cg", f(prof[42]), "x'", f(prof[43]),
              "rows", prof[44], "products", prof[45], "entries", prof[46])
    return X
U = run(Cd, Y, "user half-step")
V = run(CT, U, "item half-step")
if "--steady" in sys.argv:   # a few more epochs, then the same profile with steady-state factors
    for _ in range(3):
        U = engine.half_step(Cd, V, engine.gram(V, 0.1), algo=_lib.ALGO_TCGEN05)
        V = engine.half_step(CT, U, engine.gram(U, 0.1), algo=_lib.ALGO_TCGEN05)
    U2 = run(Cd, V, "user half-step, epoch 5")
    run(CT, U2, "item half-step, epoch 5")
