"""Host-side mirrors of the GenomicBreedingCore containers the GWAS path touches
(the Julia source is not vendored in /root/reference, field names are
taken from the call sites cited at each class).  ``missing`` is represented by NaN."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def _strs(prefix: str, n: int) -> list[str]:
    return [f"{prefix}{i + 1}" for i in range(n)]


@dataclass
class Genomes:
    """``Genomes`` (used at /root/reference/src/prediction.jl:76, :129-133): entries x
    loci-alleles allele-frequency matrix plus labels."""

    entries: list[str]
    populations: list[str]
    loci_alleles: list[str]
    allele_frequencies: np.ndarray  # n x p, Float64, NaN = missing
    mask: np.ndarray | None = None

    @classmethod
    def from_matrix(cls, A: np.ndarray, entries=None, populations=None, loci_alleles=None) -> "Genomes":
        n, p = A.shape
        return cls(
            entries=list(entries) if entries is not None else _strs("entry_", n),
            populations=list(populations) if populations is not None else ["pop_1"] * n,
            loci_alleles=list(loci_alleles) if loci_alleles is not None else _strs("chr1\t", p),
            allele_frequencies=A,
            mask=None,
        )

    def checkdims(self) -> bool:
        A = self.allele_frequencies
        if A.ndim != 2:
            return False
        n, p = A.shape
        return (len(self.entries) == n and len(self.populations) == n and len(self.loci_alleles) == p
                and len(set(self.entries)) == n and len(set(self.loci_alleles)) == p
                and (self.mask is None or self.mask.shape == A.shape))


@dataclass
class Phenomes:
    """``Phenomes`` (used at /root/reference/src/prediction.jl:114, :126)."""

    entries: list[str]
    populations: list[str]
    traits: list[str]
    phenotypes: np.ndarray  # n x t, NaN = missing
    mask: np.ndarray | None = None

    @classmethod
    def from_matrix(cls, Y: np.ndarray, entries=None, populations=None, traits=None) -> "Phenomes":
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim == 1:
            Y = Y[:, None]
        n, t = Y.shape
        return cls(
            entries=list(entries) if entries is not None else _strs("entry_", n),
            populations=list(populations) if populations is not None else ["pop_1"] * n,
            traits=list(traits) if traits is not None else _strs("trait_", t),
            phenotypes=Y,
            mask=None,
        )

    def checkdims(self) -> bool:
        Y = self.phenotypes
        if Y.ndim != 2:
            return False
        n, t = Y.shape
        return (len(self.entries) == n and len(self.populations) == n and len(self.traits) == t
                and len(set(self.entries)) == n and len(set(self.traits)) == t
                and (self.mask is None or self.mask.shape == Y.shape))


@dataclass
class Fit:
    """``Fit`` (constructed at /root/reference/src/gwas.jl:134, fields set :135-140, :231,
    :245; also /root/reference/src/linear.jl:77-98)."""

    model: str = ""
    b_hat_labels: list[str] = field(default_factory=list)
    b_hat: np.ndarray = field(default_factory=lambda: np.zeros(0))
    trait: str = ""
    entries: list[str] = field(default_factory=list)
    populations: list[str] = field(default_factory=list)
    metrics: dict = field(default_factory=dict)
    y_true: np.ndarray = field(default_factory=lambda: np.zeros(0))
    y_pred: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # extras the B200 path produces in the same pass (not in the reference's Fit)
    extras: dict = field(default_factory=dict)

    @classmethod
    def new(cls, n: int, l: int) -> "Fit":
        return cls(model="", b_hat_labels=[""] * l, b_hat=np.zeros(l), trait="", entries=[""] * n,
                   populations=[""] * n, metrics={}, y_true=np.zeros(n), y_pred=np.zeros(n))

    def checkdims(self) -> bool:
        return (len(self.b_hat_labels) == self.b_hat.shape[0] and len(self.entries) == len(self.populations)
                and len(self.entries) == self.y_true.shape[0] and self.y_true.shape[0] == self.y_pred.shape[0])


@dataclass
class GRM:
    """``GRM`` (field read at /root/reference/src/gwas.jl:121, :125)."""

    entries: list[str]
    loci_alleles: list[str]
    genomic_relationship_matrix: np.ndarray
