#!/usr/bin/env python
"""One launch of each kernel that round 2 changed or that had no ncu summary, for
`ncu --set full -k regex:...` (gpurun, one GPU):

    scan_sums_u8_tc_kernel   200,000 markers x n = 10,000 codes (2 GB)
    grm_i8_kernel            n = 10,000, p = 200,000 codes
    grm_dmma_kernel          n = 5,000, p = 100,000 Float64 (BASELINE configs[1])
    Lanczos step kernels     gbm_kstd_pc1 at n = 10,000
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import numpy as np  # noqa: E402

import gbm_b200  # noqa: E402
from gbm_b200 import _lib  # noqa: E402


def main():
    gbm_b200.init(0)
    rng = np.random.default_rng(0)
    n = 10_000
    y = rng.normal(size=n)
    pc = rng.normal(size=(n, 1))
    dm = gbm_b200.DeviceMatrix.generate(42, n, 200_000, 0)
    pk = dm.pack()
    dm.free()
    res = pk.scan(y, pc, model=1)
    print("u8 tc scan ms", _lib.last_timing()["main_ms"], "kept", int(res["keep"].sum()))
    import torch

    dK = torch.empty(n * n, dtype=torch.float64, device="cuda")
    _, tf = pk.grm(_lib.GRM_SIMPLE, 2, 0, out=dK)
    print("grm_i8 TF-equivalent", tf)
    pk.free()
    pc1, ms = gbm_b200.kstd_pc1_device(dK.data_ptr(), n)
    print("kstd_pc1 eig ms", ms)
    del dK
    g = gbm_b200.DeviceMatrix.generate(42, 5_000, 100_000, 0)
    dK = torch.empty(5_000 * 5_000, dtype=torch.float64, device="cuda")
    _, tf = g.grm(_lib.GRM_SIMPLE, 2, 0, out=dK)
    print("grm_dmma TF", tf)
    g.free()


if __name__ == "__main__":
    main()
