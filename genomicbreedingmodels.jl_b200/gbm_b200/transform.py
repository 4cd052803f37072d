"""Host-side mirror of the transformation screens, same names, keyword
arguments and error behaviour as the reference:

    square, invoneplus, log10epsdivlog10eps, mult, addnorm, raise_   /root/reference/src/transformation.jl:1-55
    transform1          /root/reference/src/transformation.jl:130-239
    transform2          /root/reference/src/transformation.jl:319-466
    epistasisfeatures   /root/reference/src/transformation.jl:540-651

The l (transform1) resp. l^2 (transform2) regressions `[1 f(x)] \\ y` run in libgbm_b200.so
(csrc/transform.cu); only labels and O(n_new) bookkeeping happen here.  `f` must be one of the
package's named endofunctions: an arbitrary closure cannot cross the C ABI, and there is no CPU
fallback, so anything else raises ArgumentError.
"""
from __future__ import annotations

from ctypes import byref, c_int64

import numpy as np

from . import _lib
from ._lib import ArgumentError, ErrorException, check, ptr
from .core import DeviceMatrix
from .gwas import _validate_and_select
from .structs import Genomes, Phenomes

_EPS = float(np.finfo(np.float64).eps)
F1_SQUARE, F1_INVONEPLUS, F1_LOG10EPS = 0, 1, 2
F2_MULT, F2_ADDNORM, F2_RAISE = 0, 1, 2


class Endofunction:
    """A named endofunction of the reference (transformation.jl:1-55): a device code plus the
    name that ends up in ``loci_alleles`` (``string(f, "(", ...)``, :235, :451)."""

    def __init__(self, name: str, arity: int, code: int):
        self.__name__ = name
        self.arity = arity
        self.code = code

    def __repr__(self):
        return self.__name__


square = Endofunction("square", 1, F1_SQUARE)
invoneplus = Endofunction("invoneplus", 1, F1_INVONEPLUS)
log10epsdivlog10eps = Endofunction("log10epsdivlog10eps", 1, F1_LOG10EPS)
mult = Endofunction("mult", 2, F2_MULT)
addnorm = Endofunction("addnorm", 2, F2_ADDNORM)
raise_ = Endofunction("raise", 2, F2_RAISE)
TRANSFORMATIONS1 = [square, invoneplus, log10epsdivlog10eps]  # defaults of epistasisfeatures (:546-547)
TRANSFORMATIONS2 = [mult, addnorm, raise_]


def _code(f, arity: int) -> int:
    if not isinstance(f, Endofunction) or f.arity != arity:
        raise ArgumentError(
            f"`{getattr(f, '__name__', f)}` is not one of the named endofunctions of {arity} argument(s) "
            "(square, invoneplus, log10epsdivlog10eps / mult, addnorm, raise): only those run on the device "
            "and there is no CPU fallback.")
    return f.code


def _extract(genomes, phenomes, idx_trait, idx_entries, idx_loci_alleles):
    """extractxyetc(...; add_intercept = false) (transformation.jl:148-156): device matrix + y + labels."""
    rows1, cols1, y = _validate_and_select(genomes, phenomes, idx_entries, idx_loci_alleles, idx_trait)
    A = np.asarray(genomes.allele_frequencies, dtype=np.float64)
    r0 = np.arange(A.shape[0]) if rows1 is None else rows1 - 1
    c0 = np.arange(A.shape[1]) if cols1 is None else cols1 - 1
    dm = DeviceMatrix.upload(A, rows1, cols1)
    entries = [genomes.entries[i] for i in r0]
    populations = [genomes.populations[i] for i in r0]
    loci_alleles = [genomes.loci_alleles[j] for j in c0]
    return dm, y, entries, populations, loci_alleles


def _check_missing(T):
    if np.isnan(T).any():
        raise ErrorException("cannot convert a value of type Missing to Float64")  # prediction.jl:129


def transform1_screen(dm: DeviceMatrix, y, f, n_new: int, eps: float = _EPS, use_abs: bool = False,
                      var_threshold: float = 0.01, want_beta: bool = True):
    """The regression loop + selection of transform1 (:165-221) on a resident matrix.  Returns
    (beta or None, idx) with idx the 1-based selected loci in sortperm order."""
    code = _code(f, 1)
    y = np.ascontiguousarray(y, dtype=np.float64)
    beta = np.empty(dm.p) if want_beta else None
    idx = np.empty(max(int(n_new), 1), dtype=np.int64)
    cnt = c_int64()
    check(_lib.lib().gbm_transform1_screen(dm._h, ptr(y), code, float(eps), int(use_abs), float(var_threshold),
                                           int(n_new), ptr(beta), ptr(idx), byref(cnt)))
    return beta, idx[:cnt.value].copy()


def transform2_screen(dm: DeviceMatrix, y, f, n_new: int, eps: float = _EPS, use_abs: bool = False,
                      var_threshold: float = 0.01, commutative: bool = False, want_beta: bool = False):
    """The pairwise loop + selection of transform2 (:362-430).  Returns (beta [l*l] or None,
    counters ascending 1-based, beta_sel)."""
    code = _code(f, 2)
    y = np.ascontiguousarray(y, dtype=np.float64)
    beta = np.empty(dm.p * dm.p) if want_beta else None
    counters = np.empty(max(int(n_new), 1), dtype=np.int64)
    vals = np.empty(max(int(n_new), 1))
    cnt = c_int64()
    check(_lib.lib().gbm_transform2_screen(dm._h, ptr(y), code, float(eps), int(use_abs), float(var_threshold),
                                           int(commutative), int(n_new), ptr(beta), ptr(counters), ptr(vals),
                                           byref(cnt)))
    return beta, counters[:cnt.value].copy(), vals[:cnt.value].copy()


def transform2_screen_rows(dm: DeviceMatrix, y, f, row0: int, row1: int, n_new: int, eps: float = _EPS,
                           use_abs: bool = False, var_threshold: float = 0.01, commutative: bool = False):
    """Rows [row0, row1) (0-based) of the pair matrix: (counters, values) of the slab's top effects in selection
    order (descending |beta|, ties by ascending global position) -- one rank's share of a sharded screen."""
    code = _code(f, 2)
    y = np.ascontiguousarray(y, dtype=np.float64)
    counters = np.empty(max(int(n_new), 1), dtype=np.int64)
    vals = np.empty(max(int(n_new), 1))
    cnt = c_int64()
    check(_lib.lib().gbm_transform2_screen_rows(dm._h, ptr(y), code, float(eps), int(use_abs), float(var_threshold),
                                                int(commutative), int(row0), int(row1), int(n_new), None,
                                                ptr(counters), ptr(vals), byref(cnt)))
    return counters[:cnt.value].copy(), vals[:cnt.value].copy()


def merge_screen_candidates(parts, n_new: int, eps: float = _EPS):
    """Merges per-shard candidate lists [(counters, values), ...] into the selection of the unsharded screen:
    sortperm(abs.(beta), rev = true)[1:n_new] is stable, so the global order is (descending |beta|, ascending
    position); keep abs(beta) > eps and return ascending positions (sort!(idx), transformation.jl:430)."""
    counters = np.concatenate([np.asarray(c, dtype=np.int64) for c, _ in parts]) if parts else np.zeros(0, np.int64)
    values = np.concatenate([np.asarray(v, dtype=np.float64) for _, v in parts]) if parts else np.zeros(0)
    order = np.lexsort((counters, -np.abs(values)))[: int(n_new)]
    order = order[np.abs(values[order]) > eps]
    asc = np.sort(counters[order])
    lookup = dict(zip(counters.tolist(), values.tolist()))
    return asc, np.array([lookup[c] for c in asc.tolist()])


def transform2_screen_sharded(dm: DeviceMatrix, y, f, n_new: int, eps: float = _EPS, use_abs: bool = False,
                              var_threshold: float = 0.01, commutative: bool = False, group=None):
    """transform2's pairwise screen over the ranks of a torch.distributed group: every rank holds the n x l
    matrix and screens a contiguous block of rows of the l x l pair matrix (no data-path collective); the
    candidate lists (<= n_new per rank) are all-gathered and merged.  Same result as transform2_screen."""
    import torch.distributed as dist

    from .sharded import shard_bounds

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    _code(f, 2)
    l = dm.p
    if n_new > l * l:  # sortperm(...)[1:n_new] on a shorter vector (transformation.jl:425); same on every rank
        raise _lib.ArgumentError(f"BoundsError: attempt to access {l * l}-element Vector{{Int64}} at index [1:{n_new}]")
    r0, r1 = shard_bounds(l, world, rank)
    # A failure on one rank only (NaN effects in its slab, out of memory) must not leave the others waiting in the
    # collective: every rank reports (error, payload), and the first error is re-raised on ALL ranks -- the same
    # ArgumentError the unsharded screen raises for a NaN effect anywhere.
    err, mine = None, (np.zeros(0, np.int64), np.zeros(0))
    try:
        if r1 > r0:
            mine = transform2_screen_rows(dm, y, f, r0, r1, n_new, eps, use_abs, var_threshold, commutative)
    except Exception as e:  # noqa: BLE001 -- re-raised below on every rank
        err = (type(e).__name__, str(e))
    parts = [(err, mine)]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (err, mine), group=group)
    for e, _ in parts:
        if e is not None:
            cls = {"ArgumentError": _lib.ArgumentError, "ErrorException": _lib.ErrorException}.get(e[0], _lib.CudaError)
            raise cls(e[1])
    return merge_screen_candidates([m for _, m in parts], n_new, eps)


def transform1_apply(dm: DeviceMatrix, f, idx, eps: float = _EPS, use_abs: bool = False) -> np.ndarray:
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    T = np.empty((dm.n, idx.size), order="F")
    check(_lib.lib().gbm_transform1_apply(dm._h, _code(f, 1), float(eps), int(use_abs), ptr(idx), idx.size, ptr(T), dm.n))
    return T


def transform2_apply(dm: DeviceMatrix, f, counters, eps: float = _EPS, use_abs: bool = False) -> np.ndarray:
    counters = np.ascontiguousarray(counters, dtype=np.int64)
    T = np.empty((dm.n, counters.size), order="F")
    check(_lib.lib().gbm_transform2_apply(dm._h, _code(f, 2), float(eps), int(use_abs), ptr(counters), counters.size,
                                          ptr(T), dm.n))
    return T


def transform1(f, genomes: Genomes, phenomes: Phenomes, idx_trait: int = 1, idx_entries=None, idx_loci_alleles=None,
               n_new_features_per_transformation: int = 1_000, ϵ: float = _EPS, use_abs: bool = False,
               σ2_threshold: float = 0.01, verbose: bool = False) -> Genomes:
    """transform1 (/root/reference/src/transformation.jl:130-239); ``σ2_threshold`` is the reference's
    ``σ²_threshold`` (not a Python identifier)."""
    _code(f, 1)
    dm, y, entries, populations, loci_alleles = _extract(genomes, phenomes, idx_trait, idx_entries, idx_loci_alleles)
    try:
        _, idx = transform1_screen(dm, y, f, n_new_features_per_transformation, ϵ, use_abs, σ2_threshold,
                                   want_beta=False)
        T = transform1_apply(dm, f, idx, ϵ, use_abs)
    finally:
        dm.free()
    _check_missing(T)
    out = Genomes(entries=entries, populations=populations,
                  loci_alleles=[f"{f.__name__}({loci_alleles[j - 1]})" for j in idx],  # :235
                  allele_frequencies=T, mask=None)
    if not out.checkdims():  # :236-238
        raise ErrorException(f"Error transforming each locus using the function `{f.__name__}`.")
    return out


def transform2(f, genomes: Genomes, phenomes: Phenomes, idx_trait: int = 1, idx_entries=None, idx_loci_alleles=None,
               n_new_features_per_transformation: int = 1_000, ϵ: float = _EPS, use_abs: bool = False,
               σ2_threshold: float = 0.01, commutative: bool = False, verbose: bool = False) -> Genomes:
    """transform2 (/root/reference/src/transformation.jl:319-466)."""
    _code(f, 2)
    dm, y, entries, populations, loci_alleles = _extract(genomes, phenomes, idx_trait, idx_entries, idx_loci_alleles)
    l = dm.p
    try:
        _, counters, _ = transform2_screen(dm, y, f, n_new_features_per_transformation, ϵ, use_abs, σ2_threshold,
                                           commutative)
        T = transform2_apply(dm, f, counters, ϵ, use_abs)
    finally:
        dm.free()
    _check_missing(T)
    names = []
    for c in counters:  # :445-451
        i, j = (c - 1) // l, (c - 1) % l
        names.append(f"{f.__name__}({loci_alleles[i]},{loci_alleles[j]})")
    out = Genomes(entries=entries, populations=populations, loci_alleles=names, allele_frequencies=T, mask=None)
    if not out.checkdims():  # :463-465
        raise ErrorException(f"Error transforming each locus using the function `{f.__name__}`.")
    return out


def epistasisfeatures(genomes: Genomes, phenomes: Phenomes, idx_trait: int = 1, idx_entries=None,
                      idx_loci_alleles=None, transformations1=None, transformations2=None,
                      n_new_features_per_transformation: int = 1_000, n_reps: int = 3, verbose: bool = False) -> Genomes:
    """epistasisfeatures (/root/reference/src/transformation.jl:540-651): n_reps rounds of every
    transformation, each round screening the features the previous ones appended."""
    t1 = list(TRANSFORMATIONS1 if transformations1 is None else transformations1)
    t2 = list(TRANSFORMATIONS2 if transformations2 is None else transformations2)
    for f in t1:
        _code(f, 1)
    for f in t2:
        _code(f, 2)
    rows1, cols1, _ = _validate_and_select(genomes, phenomes, idx_entries, idx_loci_alleles, idx_trait)  # :552-601
    A0 = np.asarray(genomes.allele_frequencies, dtype=np.float64)
    r0 = np.arange(A0.shape[0]) if idx_entries is None else np.asarray(idx_entries, dtype=np.int64) - 1
    c0 = np.arange(A0.shape[1]) if idx_loci_alleles is None else np.asarray(idx_loci_alleles, dtype=np.int64) - 1
    # slice(genomes, ...), slice(phenomes, ..., idx_traits = [idx_trait]) (:602-603)
    g = Genomes(entries=[genomes.entries[i] for i in r0], populations=[genomes.populations[i] for i in r0],
                loci_alleles=[genomes.loci_alleles[j] for j in c0],
                allele_frequencies=np.asfortranarray(A0[np.ix_(r0, c0)]), mask=None)
    ph = Phenomes(entries=[phenomes.entries[i] for i in r0], populations=[phenomes.populations[i] for i in r0],
                  traits=[phenomes.traits[idx_trait - 1]],
                  phenotypes=np.asarray(phenomes.phenotypes, dtype=np.float64)[np.ix_(r0, [idx_trait - 1])], mask=None)
    for _ in range(n_reps):  # :616
        for f in t1 + t2:
            fn = transform1 if f in t1 else transform2
            new = fn(f, g, ph, n_new_features_per_transformation=n_new_features_per_transformation)  # :618-633
            have = set(g.loci_alleles)
            cols = []
            for k, name in enumerate(new.loci_alleles):  # setdiff + first occurrence (:634-635)
                if name not in have:
                    have.add(name)
                    cols.append(k)
            g.loci_alleles = g.loci_alleles + [new.loci_alleles[k] for k in cols]  # :636-637
            g.allele_frequencies = np.asfortranarray(np.hstack([g.allele_frequencies, new.allele_frequencies[:, cols]]))
            af = g.allele_frequencies
            if af.min() < 0.0 or abs(af.max() - 1.0) > 1e-12:  # :642-650
                raise ErrorException(
                    f"The function `{f.__name__}` generates values outside the expected range of zero to one. "
                    "Please replace with an appropriate transforamtion function.")
    if not g.checkdims():  # :656-658
        raise ErrorException("Error generating new features.")
    return g
