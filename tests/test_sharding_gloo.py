"""CPU suite, part 3: the marker-sharded host logic with world_size 2 over gloo.  The
per-shard partials are made by the oracle (the checker), the plumbing under test is
gbm_b200.sharded: shard bounds, the GRM all-reduce and the shard-order gathers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    from gbm_b200.sharded import shard_bounds

    for p in (1, 7, 16, 1000, 1_000_000, 999_983):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(p, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == p
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [j1 - j0 for j0, j1 in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))
    import torch
    import torch.distributed as dist

    from gbm_b200 import sharded
    from oracle import gwas_oracle as go, synth

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, p, seed, kind = 64, 501, 9, synth.KIND_TETRAPLOID
    j0, j1 = sharded.shard_bounds(p, world, rank)
    A_loc = synth.block(seed, n, j0, j1 - j0, kind)  # this rank's column block only
    # per-shard GRM partial (what gbm_grm_accumulate produces): lower triangle of Zc Zc'
    mu = A_loc.mean(axis=0)
    Zc = A_loc - mu
    part = np.tril(Zc @ Zc.T)
    dK = torch.from_numpy(np.asfortranarray(part).T.copy().reshape(-1))  # column-major flat
    scal = torch.tensor([float(j1 - j0), float(np.sum(mu * (1 - mu)))], dtype=torch.float64)
    sharded.allreduce_grm_partials(dK, scal)
    L = dK.numpy().reshape(n, n).T
    K = (L + np.tril(L, -1).T) * sharded.grm_scale("ploidy-aware", 4, scal[0].item(), scal[1].item())
    # shard-local filter + statistics, gathered in locus order
    _, v = go.column_std(A_loc)
    idx_loc = go.fixed_locus_filter(v)
    idx = sharded.global_idx_cols(idx_loc, j0)
    y = synth.phenotype(seed, n, p, kind, n_causal=4)
    ys = (y - y.mean()) / y.std(ddof=1)
    pc = go.pca_pc1(go.standardise_K(K))
    stat_loc = np.full(j1 - j0, np.nan)
    keep = v > go.EPS
    stat_loc[keep] = go.scan_closed_form(A_loc[:, keep], ys, pc)["stat_lmm"]
    stat = sharded.gather_marker_results(stat_loc, p)
    if rank == 0:
        np.savez(os.path.join(tmp, "out.npz"), K=K, idx=idx, stat=stat[idx - 1])
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pipeline_world2_gloo(tmp_path):
    import torch.multiprocessing as mp

    from oracle import gwas_oracle as go, synth

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = np.load(tmp_path / "out.npz")
    n, p, seed, kind = 64, 501, 9, synth.KIND_TETRAPLOID
    A = synth.block(seed, n, 0, p, kind)
    y = synth.phenotype(seed, n, p, kind, n_causal=4)
    ent = [str(i) for i in range(n)]
    z, prep, _ = go.gwaslmm(A, ent, y[:, None], ent, GRM_type="ploidy-aware")
    np.testing.assert_allclose(out["K"], go.grm_ploidy_aware(A, 4), rtol=1e-11, atol=1e-13)
    assert np.array_equal(out["idx"], prep.idx_cols)  # bit-exact global filter from shard-local ones
    np.testing.assert_allclose(out["stat"], z, rtol=1e-8, atol=1e-9)


def _screen_worker(rank, world, port, tmp):
    """One rank of a sharded pairwise transformation screen: its block of rows of the l x l pair matrix comes
    from the oracle (the checker); the plumbing under test is the candidate all-gather + merge."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))
    import torch.distributed as dist

    from gbm_b200 import sharded
    from gbm_b200.transform import merge_screen_candidates
    from oracle import synth, transform_oracle as to

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, l, n_new = 40, 15, 12
    A = synth.block(3, n, 0, l, synth.KIND_TETRAPLOID)
    y = synth.phenotype(3, n, l, synth.KIND_TETRAPLOID, n_causal=3)
    r0, r1 = sharded.shard_bounds(l, world, rank)
    beta_full, _, _, _ = to.transform2(to.mult, A, y, n_new=n_new)  # rows r0..r1 are this rank's work
    slab = beta_full.reshape(l, l)[r0:r1].reshape(-1)
    order = np.argsort(-np.abs(slab), kind="stable")[:n_new]
    order = order[np.abs(slab[order]) > to.EPS]
    mine = (order + r0 * l + 1, slab[order])  # what gbm_transform2_screen_rows returns
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    counters, values = merge_screen_candidates(parts, n_new)
    if rank == 0:
        np.savez(os.path.join(tmp, "screen.npz"), counters=counters, values=values)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pairwise_screen_merge_world2_gloo(tmp_path):
    import torch.multiprocessing as mp

    from oracle import synth, transform_oracle as to

    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_screen_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = np.load(tmp_path / "screen.npz")
    n, l, n_new = 40, 15, 12
    A = synth.block(3, n, 0, l, synth.KIND_TETRAPLOID)
    y = synth.phenotype(3, n, l, synth.KIND_TETRAPLOID, n_causal=3)
    beta, idx, _, _ = to.transform2(to.mult, A, y, n_new=n_new)
    assert np.array_equal(out["counters"], idx)  # the unsharded selection, ascending (transformation.jl:430)
    assert np.array_equal(out["values"], beta[idx - 1])


def test_merge_screen_candidates_ties_and_cut():
    """Ties in |beta| keep ascending position (Julia's stable sortperm), the cut at n_new comes before the eps filter."""
    from gbm_b200.transform import merge_screen_candidates

    a = (np.array([5, 9]), np.array([2.0, -1.0]))
    b = (np.array([12, 20, 31]), np.array([-2.0, 1.0, 0.5]))
    c, v = merge_screen_candidates([b, a], 3)
    assert c.tolist() == [5, 9, 12] and v.tolist() == [2.0, -1.0, -2.0]  # |2| at 5 and 12, then |1| at 9 (before 20)
    c, v = merge_screen_candidates([a, b], 10)
    assert c.tolist() == [5, 9, 12, 20, 31]
    c, v = merge_screen_candidates([], 4)
    assert c.size == 0 and v.size == 0
