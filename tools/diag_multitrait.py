"""Device-resident multi-trait scan rate (run on the GPU box; not a test)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))
import numpy as np, torch
import gbm_b200
from gbm_b200 import _lib
gbm_b200.init(0)
n, p = 10000, 400_000
dm = gbm_b200.DeviceMatrix.generate(42, n, p, 0)
pk = dm.pack()
rng = np.random.default_rng(0)
pc = rng.normal(size=(n, 1))
for T in (1, 2, 5, 13, 20):
    Y = rng.normal(size=(n, T))
    for name, m in (("float64", dm), ("packed", pk)):
        plan = gbm_b200.ScanPlan(m, Y, pc, model=0)
        stat = torch.empty(p * T, dtype=torch.float64, device="cuda")
        for it in range(3):
            tm = plan.run(stat=stat)
        print(f"T={T:2d} {name:8s} kernel {tm['kernel_ms']:.2f} ms main(last pass) {tm['main_ms']:.2f} ms  "
              f"=> {p / tm['kernel_ms'] * 1e-3:.1f} M markers/s, {8.0 * n * p / tm['kernel_ms'] / 1e6:.0f} GB/s-equivalent of Float64", flush=True)
        plan.free()
