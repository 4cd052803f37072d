"""FP64 GRM: centred (one operand centred on the fly) against uncentred, n = 5,000 / 10,000 (GPU box; not a test)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))
import numpy as np, torch
import gbm_b200
from gbm_b200 import _lib
gbm_b200.init(0)
for n, p in ((5000, 100000), (10000, 200000)):
    dm = gbm_b200.DeviceMatrix.generate(42, n, p, 0)
    dK = torch.empty(n * n, dtype=torch.float64, device="cuda")
    for flags, name in ((0, "centred"), (_lib.GRM_NO_CENTRE, "uncentred")):
        tfs = []
        for i in range(4):
            _, tf = dm.grm(_lib.GRM_SIMPLE, 2, flags, out=dK)
            tfs.append(tf)
        print(f"n={n} p={p} {name}: {np.median(tfs[1:]):.2f} TF (SYRK credit)", flush=True)
    dm.free()
