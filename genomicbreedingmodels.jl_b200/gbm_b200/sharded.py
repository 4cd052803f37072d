"""Marker-sharded multi-GPU host logic: one process per GPU, `torch.distributed` for the
plumbing.

The reference's only parallelism is `Threads.@threads` over markers
(/root/reference/src/gwas.jl:239, :363); the B200 equivalent is a contiguous column block
per GPU.  The scan needs no collective -- results are gathered in shard order -- and the
GRM is the sum of per-shard partials: ONE all-reduce of the n x n lower-triangle partial
plus two scalars (NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors in
the tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(p: int, world_size: int, rank: int) -> tuple[int, int]:
    """Columns [j0, j1) of rank `rank`: [p*r/W, p*(r+1)/W) (0-based, contiguous, ascending)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return (p * rank) // world_size, (p * (rank + 1)) // world_size


def allreduce_grm_partials(dK, scalars, group=None):
    """Sum the per-shard GRM partials in place.  dK: n*n float64 tensor (lower triangle of
    sum_j (a_j-mu_j)(a_j-mu_j)' over the shard), scalars: float64 tensor
    [p_shard, sum_j q_j(1-q_j)].  Tensors on CUDA use NCCL, on CPU gloo."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(dK, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(scalars, op=dist.ReduceOp.SUM, group=group)
    return dK, scalars


def grm_scale(grm_type: str, ploidy: int | None, p_total: float, sum_q1mq: float) -> float:
    """simple: 1/p ; ploidy-aware: ploidy / sum_j q_j (1 - q_j)."""
    if grm_type == "ploidy-aware":
        if not sum_q1mq > 0:
            raise ValueError("sum q(1-q) is not positive")
        return float(ploidy) / float(sum_q1mq)
    return 1.0 / float(p_total)


def gather_marker_results(local: np.ndarray, p_total: int, group=None):
    """Concatenate per-shard per-marker arrays in shard (= locus) order on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == p_total
    _ = torch
    return out


def global_idx_cols(local_idx_cols: np.ndarray, j0: int, group=None) -> np.ndarray:
    """idx_cols of the whole problem (1-based, ascending) from per-shard 1-based indices:
    shard-order concatenation with the shard's column offset added."""
    import torch.distributed as dist

    shifted = np.asarray(local_idx_cols, dtype=np.int64) + j0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return shifted
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, shifted, group=group)
    return np.concatenate(parts)


class ShardedGWAS:
    """gwasols / gwaslmm over a column-sharded device matrix (one shard per rank).

    `dm` is this rank's DeviceMatrix holding columns [j0, j1) of the n x p_total problem."""

    def __init__(self, dm, p_total: int, j0: int, group=None):
        self.dm, self.p_total, self.j0, self.group = dm, int(p_total), int(j0), group

    def grm(self, grm_type: str = "simple", ploidy: int = 2):
        """All ranks end with the full symmetric GRM as a CUDA tensor (n*n, column-major)."""
        import torch

        from . import core

        n = self.dm.n
        dK = torch.zeros(n * n, dtype=torch.float64, device="cuda")
        s, _ = self.dm.grm_accumulate(dK.data_ptr(), centre=True)
        scal = torch.tensor([float(self.dm.p), s], dtype=torch.float64, device="cuda")
        allreduce_grm_partials(dK, scal, self.group)
        p_tot, sq = (float(x) for x in scal.cpu())
        core.grm_finalize(dK.data_ptr(), n, grm_scale(grm_type, ploidy, p_tot, sq))
        return dK

    def scan(self, ys, pc1, model: int):
        res = self.dm.scan(ys, pc1[:, None], model=model)
        st = self.dm.colstats()
        idx = global_idx_cols(st["idx_cols"], self.j0, self.group)
        stat = gather_marker_results(res["stat"][:, 0], self.p_total, self.group)
        return stat[idx - 1], idx
