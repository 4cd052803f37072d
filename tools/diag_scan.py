import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))
import numpy as np, torch
import gbm_b200
from gbm_b200 import _lib
gbm_b200.init(0)
n, p = 10000, 1_000_000
dm = gbm_b200.DeviceMatrix.generate(42, n, p, 0)
rng = np.random.default_rng(0); y = rng.normal(size=n)
dK = torch.empty(n * n, dtype=torch.float64, device="cuda")
for it in range(2):
    t0 = time.perf_counter(); dm.grm(0, 2, 0, out=dK); t1 = time.perf_counter()
    pc, eig = gbm_b200.kstd_pc1_device(dK.data_ptr(), n); t2 = time.perf_counter()
    res = dm.scan(y, pc[:, None], model=1); t3 = time.perf_counter()
    tm = _lib.last_timing()
    res = dm.scan(y, pc[:, None], model=1); t4 = time.perf_counter()
    print(f"grm {t1-t0:.3f} pc1 {t2-t1:.3f} scan {t3-t2:.4f} scan-again {t4-t3:.4f}", tm, flush=True)
    print("pc finite", np.isfinite(pc).all(), "nan stats", np.isnan(res["stat"]).sum(), flush=True)
