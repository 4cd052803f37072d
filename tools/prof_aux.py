"""Auxiliary profiling driver (run under ncu on the GPU box): K standardisation + PC1 at
n = 10,000 and the packed (1-byte) scan.  Not part of the tests or the bench."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import numpy as np
import torch

import gbm_b200
from gbm_b200 import _lib

gbm_b200.init(0)
n, p = 10000, 60000
dm = gbm_b200.DeviceMatrix.generate(42, n, p, 0)
dK = torch.empty(n * n, dtype=torch.float64, device="cuda")
dm.grm(_lib.GRM_SIMPLE, 2, 0, out=dK)
for it in range(2):
    t0 = time.perf_counter()
    pc, eig_ms = gbm_b200.kstd_pc1_device(dK.data_ptr(), n)
    print(f"kstd_pc1 n={n}: total {time.perf_counter() - t0:.3f} s, cuSOLVER {eig_ms * 1e-3:.3f} s", flush=True)
pk = dm.pack()
rng = np.random.default_rng(0)
y, c = rng.normal(size=n), rng.normal(size=n)
plan = gbm_b200.ScanPlan(pk, y, c[:, None], model=1)
stat = torch.empty(p, dtype=torch.float64, device="cuda")
for it in range(3):
    tm = plan.run(stat=stat)
print("packed scan", tm, flush=True)
