"""Multi-GPU check, one process per GPU under torchrun on N GPUs of one box; collected by pytest through
tests/test_gpu_group.py::test_one_process_per_gpu_under_torchrun (marker `multigpu`), or by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_gpu_check.py

Marker-sharded gwaslmm through the library's RANK group (gbm_group_create_rank: NCCL inside libgbm_b200.so,
torch.distributed only hands out the 128-byte id): per-rank column block generated on the device, GRM partials
summed with one all-reduce, PC1 with the columns of K sharded, scan per rank, results gathered in locus order
on every rank and compared on rank 0 with the single-process oracle; then the one-call gbm_sharded_gwas."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import numpy as np
import torch
import torch.distributed as dist

import gbm_b200
from gbm_b200 import _lib, sharded
from oracle import gwas_oracle as go, synth


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gbm_b200.init(local)
    n, p, seed, kind = 512, 6000, 17, synth.KIND_TETRAPLOID
    j0, j1 = sharded.shard_bounds(p, world, rank)
    dm = gbm_b200.DeviceMatrix.generate(seed, n, j1 - j0, kind, col0=j0)
    sg = sharded.ShardedGWAS(dm, p, j0)
    K = sg.grm("ploidy-aware", ploidy=4)
    pc = sg.pc1()
    y = synth.phenotype(seed, n, p, kind)
    ys = (y - y.mean()) / y.std(ddof=1)
    z, idx = sg.scan(ys, pc, _lib.MODEL_LMM)
    one = sg.sm.gwas(ys, model=_lib.MODEL_LMM, grm_type=_lib.GRM_PLOIDY_AWARE)  # the whole pipeline in ONE call
    # sharded PC1 at a size where the Lanczos path runs (n >= 1024): against the single-GPU routine on rank 0's GPU
    n2, p2 = 4224, 4096
    c0, c1 = sharded.shard_bounds(p2, world, rank)
    dm2 = gbm_b200.DeviceMatrix.generate(seed, n2, c1 - c0, synth.KIND_DIPLOID, col0=c0)
    sg2 = sharded.ShardedGWAS(dm2, p2, c0)
    K2 = sg2.grm("simple")
    pc2 = sg2.pc1()                       # per-step all-reduce over peer memory (CUDA IPC mailboxes, peer_sum_kernel)
    t_peer = time.perf_counter()
    pc2 = sg2.pc1()
    t_peer = time.perf_counter() - t_peer
    os.environ["GBM_PC1_PEER"] = "0"      # the same through NCCL
    pc2_nccl = sg2.pc1()
    t_nccl = time.perf_counter()
    pc2_nccl = sg2.pc1()
    t_nccl = time.perf_counter() - t_nccl
    del os.environ["GBM_PC1_PEER"]
    # larger GRM timing incl. the all-reduce
    big_n, big_p = 4096, 65536 * world
    b0, b1 = sharded.shard_bounds(big_p, world, rank)
    bm = gbm_b200.DeviceMatrix.generate(seed, big_n, b1 - b0, synth.KIND_DIPLOID, col0=b0)
    sb = sharded.ShardedGWAS(bm, big_p, b0)
    sb.grm("simple", want_host=False)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    sb.grm("simple", want_host=False)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        A = synth.block(seed, n, 0, p, kind)
        ent = [str(i) for i in range(n)]
        z_ref, prep, _ = go.gwaslmm(A, ent, y[:, None], ent, GRM_type="ploidy-aware")
        Kref = go.grm_ploidy_aware(A, 4)
        e_k = np.max(np.abs(K - Kref)) / np.abs(Kref).max()
        e_z = np.max(np.abs(z - z_ref) / np.maximum(np.abs(z_ref), 1e-3 * np.abs(z_ref).max()))
        z1 = one["stat"][one["idx_cols"] - 1]
        e_1 = np.max(np.abs(z1 - z_ref) / np.maximum(np.abs(z_ref), 1e-3 * np.abs(z_ref).max()))
        _, pc2_ref, _ = gbm_b200.kstd_pc1(K2, want_kstd=False)
        e_pc = min(np.max(np.abs(pc2 - pc2_ref)), np.max(np.abs(pc2 + pc2_ref)))
        e_pc = max(e_pc, min(np.max(np.abs(pc2_nccl - pc2_ref)), np.max(np.abs(pc2_nccl + pc2_ref))))
        ok = (np.array_equal(idx, prep.idx_cols) and np.array_equal(one["idx_cols"], prep.idx_cols) and e_k < 1e-11
              and e_z < 1e-9 and e_1 < 1e-9 and one["timing"]["ploidy"] == 4 and e_pc < 1e-9)
        tf = big_n * (big_n + 1) * big_p / dt / 1e12
        print(f"dist_gpu_check world={world}: idx_cols equal={np.array_equal(idx, prep.idx_cols)} "
              f"grm rel err={e_k:.2e} z rel err={e_z:.2e} one-call z rel err={e_1:.2e} "
              f"sharded-Lanczos PC1 (n={n2}) vs single-GPU {e_pc:.2e} (peer memory {t_peer*1e3:.1f} ms, NCCL {t_nccl*1e3:.1f} ms) | sharded GRM n={big_n} p={big_p}: {dt*1e3:.1f} ms "
              f"= {tf:.1f} TFLOP/s aggregate incl. all-reduce -> {'OK' if ok else 'FAIL'}", flush=True)
        if not ok:
            sys.exit(1)
    # pairwise transformation screen, rows of the pair matrix sharded over the ranks (every rank holds the matrix)
    from gbm_b200 import transform as tr

    tl, tn_new = 700, 300
    tm = gbm_b200.DeviceMatrix.generate(seed, 800, tl, synth.KIND_TETRAPLOID)
    ty = synth.phenotype(seed, 800, tl, synth.KIND_TETRAPLOID)
    sc, sv = tr.transform2_screen_sharded(tm, ty, tr.mult, tn_new)
    if rank == 0:
        _, uc, uv = tr.transform2_screen(tm, ty, tr.mult, tn_new)
        same = np.array_equal(sc, uc) and np.array_equal(sv, uv)
        print(f"dist_gpu_check world={world}: sharded transform2 screen (l={tl}, {tl * tl} regressions) "
              f"selection identical to the unsharded one: {same} -> {'OK' if same else 'FAIL'}", flush=True)
        if not same:
            sys.exit(1)
    tm.free()
    for g in (sg, sg2, sb):
        g.free()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
