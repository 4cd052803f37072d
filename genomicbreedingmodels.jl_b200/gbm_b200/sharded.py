"""Marker-sharded host logic: shard bounds, the reduction / gather layout and scale rules.

On the GPU all of this runs INSIDE libgbm_b200.so (csrc/group.cu, `gbm_group` / `gbm_sharded`, NCCL); the
functions below restate that layout over `torch.distributed` tensors so that it can be exercised with
world_size 2 over gloo on a CPU-only box (tests/test_sharding_gloo.py), and `ShardedGWAS` is the
one-process-per-GPU convenience wrapper over the library's rank group.

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


def _gather_variable(local, group=None):
    """All-gather of per-rank 1-D arrays of different lengths as TENSORS (no pickling): counts first, then
    equal-sized padded blocks.  Backend-agnostic host logic (gloo in the CPU tests); on the GPU the library
    does the same inside gbm_sharded_* over NCCL (csrc/group.cu: gather_rows)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local))
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([t.numel()], dtype=torch.int64), group=group)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    padded = torch.zeros(width, dtype=t.dtype)
    padded[: t.numel()] = t
    blocks = [torch.empty(width, dtype=t.dtype) for _ in range(world)]
    dist.all_gather(blocks, padded, group=group)
    return np.concatenate([b[:c].numpy() for b, c in zip(blocks, counts)])


def gather_marker_results(local: np.ndarray, p_total: int, group=None):
    """Concatenate per-shard per-marker arrays in shard (= locus) order on every rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    out = _gather_variable(local, group)
    assert out.shape[0] == p_total
    return out


def global_idx_cols(local_idx_cols: np.ndarray, j0: int, group=None) -> np.ndarray:
    """idx_cols of the whole problem (1-based, ascending) from per-shard 1-based indices:
    shard-order concatenation with the shard's column offset added."""
    import torch.distributed as dist

    shifted = np.asarray(local_idx_cols, dtype=np.int64) + j0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return shifted
    return _gather_variable(shifted, group)


class ShardedGWAS:
    """gwasols / gwaslmm over a column-sharded device matrix, one process per GPU (torchrun): a thin wrapper
    over the library's rank group (``gbm_group_create_rank`` + ``gbm_sharded_adopt``).  All collectives run inside
    libgbm_b200.so over NCCL; torch.distributed only hands out the 128-byte group id.

    `dm` is this rank's DeviceMatrix holding its column block of the n x p_total problem."""

    def __init__(self, dm, p_total: int, j0: int, group=None):
        from . import multigpu

        self.grp = multigpu.Group.from_torch_distributed(group)
        self.sm = multigpu.ShardedMatrix.adopt(self.grp, [dm])
        if self.sm.p != int(p_total) or self.sm.first_col[0] != int(j0):
            raise ValueError("the ranks' blocks do not tile the matrix in rank order")
        self.dm, self.p_total, self.j0 = dm, int(p_total), int(j0)

    def grm(self, grm_type: str = "simple", ploidy: int = 2, want_host: bool = True):
        """Full symmetric GRM (host copy on every rank when ``want_host``); it also stays resident for ``pc1``."""
        from . import _lib

        code = _lib.GRM_PLOIDY_AWARE if grm_type == "ploidy-aware" else _lib.GRM_SIMPLE
        K, _ = self.sm.grm(code, ploidy, 0, want_host=want_host)
        return K

    def pc1(self):
        return self.sm.kstd_pc1()[0]

    def scan(self, ys, pc1, model: int):
        res = self.sm.scan(ys, pc1[:, None], model=model)
        idx = self.sm.colstats()["idx_cols"]
        return res["stat"][idx - 1, 0], idx

    def free(self):
        self.sm.free()
        self.grp.free()
