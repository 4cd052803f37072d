"""Where does the time of DeviceMatrix.scan go at full size?  (run on the GPU box; not a test)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))
import numpy as np, torch
import gbm_b200
from gbm_b200 import _lib
gbm_b200.init(0)
n, p = 10000, 1_000_000
dm = gbm_b200.DeviceMatrix.generate(42, n, p, 0)
pk = dm.pack()
rng = np.random.default_rng(0); y = rng.normal(size=n); pc = rng.normal(size=n)
for name, m in (("float64", dm), ("packed", pk)):
    for it in range(4):
        t0 = time.perf_counter(); res = m.scan(y, pc[:, None], model=1); t1 = time.perf_counter()
        print(name, it, f"scan {t1-t0:.4f}", _lib.last_timing(), flush=True)
    plan = gbm_b200.ScanPlan(m, y, pc[:, None], model=1)
    outs = {k: np.empty(p) for k in ("beta", "se", "stat", "nlp", "mean", "sd")}
    keep = np.empty(p, dtype=np.uint8)
    for it in range(3):
        t0 = time.perf_counter(); tm = plan.run(outs["beta"], outs["se"], outs["stat"], outs["nlp"], outs["mean"], outs["sd"], keep); t1 = time.perf_counter()
        print(name, "plan.run host outputs", f"{t1-t0:.4f}", tm, flush=True)
    plan.free()
