"""PC1 solver timing at n = 10,000 (run on the GPU box; not a test)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))
import numpy as np, torch
import gbm_b200
from gbm_b200 import _lib
gbm_b200.init(0)
for n, p in ((10000, 200000), (20000, 100000)):
    dm = gbm_b200.DeviceMatrix.generate(42, n, p, 0)
    pk = dm.pack()
    dm.free()
    dK = torch.empty(n * n, dtype=torch.float64, device="cuda")
    pk.grm(0, 2, 0, out=dK)
    for solver in ("lanczos", "lanczos-gram", "cusolver"):
        os.environ["GBM_PC1_SOLVER"] = solver
        for it in range(2):
            t0 = time.perf_counter()
            pc, eig_ms = gbm_b200.kstd_pc1_device(dK.data_ptr(), n)
            dt = time.perf_counter() - t0
        print(f"n={n} {solver}: kstd_pc1 {dt:.3f} s, eig {eig_ms:.1f} ms, launches {_lib.last_timing()['launches']}", flush=True)
    pk.free()
    del dK
