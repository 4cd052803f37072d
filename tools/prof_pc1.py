"""One gbm_kstd_pc1 per reorthogonalisation route at n = 10,000 on one GPU, for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"reorth|dots_kernel|project_out|norm_next|partial_reduce|gram_fused" \
        --csv --log-file gpurun_out/pc1_launches.csv python tools/prof_pc1.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import torch  # noqa: E402

import gbm_b200  # noqa: E402
from gbm_b200 import _lib  # noqa: E402

gbm_b200.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
dm = gbm_b200.DeviceMatrix.generate(42, n, 200_000, 0)
pk = dm.pack()
dm.free()
dK = torch.empty(n * n, dtype=torch.float64, device="cuda")
pk.grm(_lib.GRM_SIMPLE, 2, 0, out=dK)
pk.free()
for no_coop in ("1", "0", "1", "0"):
    os.environ["GBM_PC1_NO_COOP"] = no_coop
    pc, eig_ms = gbm_b200.kstd_pc1_device(dK.data_ptr(), n)
    print(f"n={n} GBM_PC1_NO_COOP={no_coop}: eig {eig_ms:.2f} ms, steps ~{_lib.last_timing()['launches']}", flush=True)
