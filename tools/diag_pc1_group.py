"""PC1 (K standardisation + Lanczos) timing at n = 10,000 on a local group of all visible GPUs and on one GPU, every
route: cooperative reorthogonalisation on / off (GBM_PC1_NO_COOP), per-step all-reduce over peer memory / NCCL
(GBM_PC1_PEER).  Run on the GPU box (not a test):  python tools/diag_pc1_group.py [n_gpus] [n]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import numpy as np  # noqa: E402

import gbm_b200  # noqa: E402
from gbm_b200 import _lib, multigpu  # noqa: E402


def main():
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    gbm_b200.init(0)
    ref = None
    for world in sorted({1, n_gpus}):
        grp = multigpu.Group.local(world)
        sm = multigpu.ShardedMatrix.generate(grp, 42, n, 131_072, 0, pack=True)  # the same matrix for every group size
        sm.grm(_lib.GRM_SIMPLE, want_host=False)
        for coop in ("0", "1"):
            for peer in (("1", "0") if world > 1 else ("-",)):
                os.environ["GBM_PC1_NO_COOP"] = "0" if coop == "1" else "1"
                if peer != "-":
                    os.environ["GBM_PC1_PEER"] = peer
                best = None
                for _ in range(2):
                    t0 = time.perf_counter()
                    pc, eig_ms = sm.kstd_pc1()
                    dt = (time.perf_counter() - t0) * 1e3
                    best = (dt, eig_ms) if best is None or dt < best[0] else best
                if ref is None:
                    ref = pc
                err = min(np.max(np.abs(pc - ref)), np.max(np.abs(pc + ref)))
                print(f"n={n} GPUs={world} cooperative-reorth={coop} peer-allreduce={peer}: kstd_pc1 {best[0]:.1f} ms, "
                      f"eig {best[1]:.1f} ms, max |pc - first route's pc| {err:.1e}", flush=True)
        sm.free()
        grp.free()


if __name__ == "__main__":
    main()
