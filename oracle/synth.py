"""Counter-based synthetic allele-frequency generator (TEST INFRASTRUCTURE / data definition).

This is the NumPy definition of the synthetic genotype matrix used by the parity tests
and by bench.py.  The CUDA generator in
``genomicbreedingmodels.jl_b200/csrc/generate.cu`` implements the same integer
arithmetic, so any column block of a device-generated matrix can be regenerated on the
CPU bit-exactly without ever holding the whole matrix on the host.

The reference's own simulators (``GenomicBreedingCore.simulategenomes`` /
``simulatetrials``, call sites /root/reference/src/gwas.jl:41-51) are Julia code that
is absent from this image; the shapes they produce (entries x loci-alleles, values in
[0,1], ploidy-level dosages after ``round.(af .* ploidy) ./ ploidy``,
/root/reference/src/gwas.jl:43-45) are what this generator imitates.

All arithmetic that decides a value is 64-bit integer arithmetic; every emitted double
is a dyadic rational (k/2, k/4 or k/4096), hence exact in Float64.
"""
from __future__ import annotations

import numpy as np

KIND_DIPLOID = 0  # Binomial(2, q_j) / 2  (BASELINE configs C2, C3)
KIND_TETRAPLOID = 1  # Binomial(4, q_j) / 4  (C4; matches the doctests' rounding)
KIND_CONTINUOUS = 2  # q_j + noise, on a 1/4096 grid, clipped to [0,1]  (C1)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)
_COLSALT = np.uint64(0xD1B54A32D192ED03)
_FIXSALT = np.uint64(0x8CB92BA72F3D8DD7)

# one column in FIXED_ONE_IN is constant (exercises the fixed-locus filter,
# /root/reference/src/gwas.jl:112-115)
FIXED_ONE_IN = 97


def mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _C1
        z = (z ^ (z >> np.uint64(27))) * _C2
        z = z ^ (z >> np.uint64(31))
    return z


def _col_hash(seed: int, j):
    j = np.asarray(j, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed) * _GOLD + (j + np.uint64(1)) * _COLSALT)


def column_params(seed: int, j):
    """Per-column integer parameters: (thr16, fixed_flag, fixed_level16).

    thr16  : allele-frequency threshold on a 16-bit grid, q_j = thr16/65536 in [0.05, 0.5)
    fixed  : True when the column is constant
    level  : for fixed columns, selects which constant (0, 1 or a mid level)
    """
    h = _col_hash(seed, j)
    u16 = (h >> np.uint64(48)).astype(np.uint64)  # 16 bits
    thr = np.uint64(3277) + ((u16 * np.uint64(29491)) >> np.uint64(16))
    hf = mix64(h ^ _FIXSALT)
    fixed = (hf % np.uint64(FIXED_ONE_IN)) == np.uint64(0)
    level = (hf >> np.uint64(32)) % np.uint64(3)
    return thr, fixed, level


def block(seed: int, n: int, j0: int, ncols: int, kind: int, rows=None) -> np.ndarray:
    """Columns j0 .. j0+ncols-1 (0-based, global column index) as an (n, ncols)
    Fortran-ordered Float64 array.  ``rows`` (optional 0-based row indices) restricts the
    output to those rows, (len(rows), ncols): values depend on (seed, row, column) only."""
    i = (np.arange(n, dtype=np.uint64) if rows is None else np.asarray(rows, dtype=np.uint64))[:, None]
    j = (np.arange(ncols, dtype=np.uint64) + np.uint64(j0))[None, :]
    thr, fixed, level = column_params(seed, j)
    hcol = _col_hash(seed, j)
    with np.errstate(over="ignore"):
        h = mix64(hcol + (i + np.uint64(1)) * _GOLD)
    f = [(h >> np.uint64(16 * k)) & np.uint64(0xFFFF) for k in range(4)]
    if kind == KIND_DIPLOID:
        dos = (f[0] < thr).astype(np.int64) + (f[1] < thr).astype(np.int64)
        a = dos.astype(np.float64) * 0.5
        fixed_val = np.where(level == 0, 0.0, np.where(level == 1, 1.0, 0.5))
    elif kind == KIND_TETRAPLOID:
        dos = sum((fk < thr).astype(np.int64) for fk in f)
        a = dos.astype(np.float64) * 0.25
        fixed_val = np.where(level == 0, 0.0, np.where(level == 1, 1.0, 0.25))
    elif kind == KIND_CONTINUOUS:
        # q on a 1/4096 grid plus a centred triangular noise of +-1024/4096
        base = (thr >> np.uint64(4)).astype(np.int64)  # thr/16 -> [204, 2048)
        noise = ((f[0] >> np.uint64(6)).astype(np.int64) + (f[1] >> np.uint64(6)).astype(np.int64)) - 1024
        m = np.clip(base + noise, 0, 4096)
        a = m.astype(np.float64) * (1.0 / 4096.0)
        # a non-dyadic constant for level 2: exercises the two-pass/shifted variance
        fixed_val = np.where(level == 0, 0.0, np.where(level == 1, 1.0, 0.3))
    else:
        raise ValueError("unknown kind")
    a = np.where(fixed, np.broadcast_to(fixed_val, a.shape), a)
    return np.asfortranarray(a)


def phenotype(seed: int, n: int, p: int, kind: int, n_causal: int = 32, h2: float = 0.5) -> np.ndarray:
    """Synthetic trait y = sum of n_causal evenly spaced marker effects + noise with
    heritability h2 (cf. ``f_add_dom_epi`` / ``proportion_of_variance``,
    /root/reference/src/gwas.jl:47-49).  Only the causal columns are generated, so
    this is O(n * n_causal) whatever p is.  Host-side only (the device never makes y)."""
    rng = np.random.default_rng(seed ^ 0x5EED)
    n_causal = max(1, min(n_causal, p))
    causal = np.unique((np.arange(n_causal, dtype=np.int64) * p) // n_causal + (p // (2 * n_causal)))
    causal = causal[causal < p]
    g = np.zeros(n)
    beta = rng.normal(size=causal.size)
    for b, j in zip(beta, causal):
        g += b * block(seed, n, int(j), 1, kind)[:, 0]
    vg = g.var()
    e = rng.normal(size=n)
    if vg > 0:
        e *= np.sqrt(vg * (1.0 - h2) / h2) / e.std()
    return g + e
