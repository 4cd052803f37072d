"""GRM-covariance LMM engine once, for `ncu --set full -k regex:lmm_delta` (gpurun, one GPU):
n = 8,000, 32,768 markers: GRM, syevd (cuSOLVER), rotation GEMM, per-marker REML delta search."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gbm_b200  # noqa: E402
from gbm_b200 import _lib  # noqa: E402

gbm_b200.init(0)
n, pm = 8_000, 32_768
dm = gbm_b200.DeviceMatrix.generate(42, n, pm, 0)
dK = torch.empty(n * n, dtype=torch.float64, device="cuda")
dm.grm(_lib.GRM_SIMPLE, 2, 0, out=dK)
rng = np.random.default_rng(3)
g = np.zeros(n)
for j in rng.choice(pm, size=500, replace=False):
    c = dm.download(int(j), 1)[:, 0]
    g += rng.normal() * (c - c.mean())
g /= g.std()
y = np.sqrt(0.5) * g + np.sqrt(0.5) * rng.normal(size=n)
plan = gbm_b200.LmmPlan(dK, y)
for flags, name in ((0, "reml"), (_lib.LMM_REFERENCE_OBJECTIVE, "reference objective")):
    res = plan.run(dm, flags=flags)
    print(f"{name}: n={n} markers={pm}: rotation {res['timing']['main_ms']:.1f} ms = {res['gemm_tflops']:.1f} TF, "
          f"delta search {res['search_ms']:.1f} ms, log delta in [{np.nanmin(res['log_delta']):.2f}, {np.nanmax(res['log_delta']):.2f}]",
          flush=True)
plan.free()
dm.free()
