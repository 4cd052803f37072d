"""CPU oracle for the transformation screens -- TEST INFRASTRUCTURE ONLY.

Literal NumPy restatement of /root/reference/src/transformation.jl:
    named endofunctions            :1-55     (square, invoneplus, log10epsdivlog10eps, mult, addnorm, raise)
    transform1                     :130-239  (per-locus OLS screen of f(x))
    transform2                     :319-466  (pairwise OLS screen of f(x_i, x_j))
    epistasisfeatures              :540-651  (n_reps rounds of both, new features appended)
and of the regression they call, ``ols`` -> ``b_hat = X \\ y`` with X = [1 f(x)]
(/root/reference/src/linear.jl:68-85), whose second coefficient is the screen statistic.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; nothing under
genomicbreedingmodels.jl_b200/ does.

Pinning: the reference's doctests for this path pin that a selected feature equals f applied to
its source locus (transformation.jl:113-126, :298-316) and the output ranges of
epistasisfeatures (:522-536); tests/test_transform_oracle.py re-asserts those.  They hold no numeric
beta vector, and Julia is not available in this image, so the beta values are PARITY UNPINNED
against the reference itself; they are checked here against the closed form of the 2-column
least-squares problem.

``X \\ y`` for a tall Matrix{Float64} is Julia's pivoted QR with rank truncation at
rcond = min(size(X)) * eps = 2 eps on the singular-value ratio estimate (LinearAlgebra
``ldiv!(::QRPivoted, B, rcond)``, the xGELSY scheme) and the minimum-norm solution when X is
rank deficient -- e.g. addnorm of two complementary alleles is exactly constant.
``numpy.linalg.lstsq(X, y, rcond=2 eps)`` applies the same criterion to the singular values and
returns the same minimum-norm solution.
"""
from __future__ import annotations

import numpy as np

EPS = float(np.finfo(np.float64).eps)


# ---- named endofunctions (transformation.jl:9, :18, :27, :36, :45, :54) ----------------
def square(x):
    return x * x  # Julia lowers x^2 to x*x


def invoneplus(x):
    return 1.0 / (1.0 + x)


def log10epsdivlog10eps(x):
    return np.log10(x + EPS) / np.log10(EPS)


def mult(x, y):
    return x * y


def addnorm(x, y):
    return (x + y) / 2.0


def raise_(x, y):
    return np.power(x, y)


raise_.__name__ = "raise"
TRANSFORMATIONS1 = [square, invoneplus, log10epsdivlog10eps]  # transformation.jl:546
TRANSFORMATIONS2 = [mult, addnorm, raise_]  # :547


def ols_slope(z: np.ndarray, y: np.ndarray) -> float:
    """fit.b_hat[2] of ols(genomes = g, phenomes = p) with one locus (transformation.jl:201-202,
    linear.jl:85)."""
    X = np.column_stack([np.ones(z.shape[0]), z])
    b = np.linalg.lstsq(X, y, rcond=2.0 * EPS)[0]
    return float(b[1])


def _prep(X, eps, use_abs):
    X = np.asarray(X, dtype=np.float64) + eps  # :158 / :347
    if use_abs:
        X = np.abs(X)  # :160-162
    return X


def _select(beta, n_new, eps):
    """sortperm(abs.(beta), rev = true)[1:n_new], then keep abs(beta) > eps (:212-221, :420-429).
    Julia's sortperm is stable, also with rev = true."""
    if n_new > beta.shape[0]:
        raise IndexError("BoundsError: attempt to access %d-element Vector at index [1:%d]" % (beta.shape[0], n_new))
    order = np.argsort(-np.abs(beta), kind="stable")[:n_new]
    return np.array([j for j in order if abs(beta[j]) > eps], dtype=np.int64)


def _clean(T, eps):
    T = T.copy()
    T[np.abs(T) < eps] = 0.0  # :224-225
    T[np.abs(T - 1.0) < eps] = 1.0  # :226-227
    return T


def transform1(f, X, y, n_new=1000, eps=EPS, use_abs=False, var_threshold=0.01):
    """Returns (beta [l], idx 1-based in selection order, T [n x len(idx)])."""
    X = _prep(X, eps, use_abs)
    n, l = X.shape
    beta = np.zeros(l)
    for j in range(l):
        x = X[:, j]
        if np.var(x, ddof=1) < var_threshold:  # :183
            continue
        beta[j] = ols_slope(f(x), y)  # :189-202
    idx = _select(beta, n_new, eps)
    T = _clean(f(X[:, idx]), eps) if idx.size else np.zeros((n, 0))
    return beta, idx + 1, T


def transform2(f, X, y, n_new=1000, eps=EPS, use_abs=False, var_threshold=0.01, commutative=False):
    """Returns (beta [l*l], counters 1-based ascending, pairs (i, j) 1-based, T)."""
    X = _prep(X, eps, use_abs)
    n, l = X.shape
    beta = np.zeros(l * l)
    v = np.var(X, axis=0, ddof=1)
    counter = 0
    for i in range(l):
        for j in range(l):
            counter += 1  # :372
            if commutative and j < i:  # :373
                continue
            if v[i] < var_threshold or v[j] < var_threshold:  # :381
                continue
            beta[counter - 1] = ols_slope(f(X[:, i], X[:, j]), y)
    idx = np.sort(_select(beta, n_new, eps))  # :430 sort!(idx)
    pairs = [(int(c // l) + 1, int(c % l) + 1) for c in idx]  # :445-449 (c is zero-based here)
    T = np.zeros((n, idx.size))
    for k, (i, j) in enumerate(pairs):
        T[:, k] = f(X[:, i - 1], X[:, j - 1])
    return beta, idx + 1, pairs, _clean(T, eps)


def feature_name(f, *loci):
    return f.__name__ + "(" + ",".join(loci) + ")"  # :235, :451


def epistasisfeatures(A, y, loci_alleles, transformations1=None, transformations2=None, n_new=1000, n_reps=3):
    """epistasisfeatures (:616-651): returns (allele_frequencies, loci_alleles) with the new features appended."""
    t1 = TRANSFORMATIONS1 if transformations1 is None else transformations1
    t2 = TRANSFORMATIONS2 if transformations2 is None else transformations2
    A = np.array(A, dtype=np.float64)
    names = list(loci_alleles)
    for _ in range(n_reps):
        for f in list(t1) + list(t2):
            if f in t1:
                _, idx, T = transform1(f, A, y, n_new=n_new)
                new_names = [feature_name(f, names[j - 1]) for j in idx]
            else:
                _, _, pairs, T = transform2(f, A, y, n_new=n_new)
                new_names = [feature_name(f, names[i - 1], names[j - 1]) for i, j in pairs]
            have = set(names)
            cols = []
            for k, nm in enumerate(new_names):  # setdiff + first occurrence (:641-643)
                if nm not in have:
                    have.add(nm)
                    cols.append(k)
            names += [new_names[k] for k in cols]
            A = np.hstack([A, T[:, cols]])
            if A.min() < 0.0 or abs(A.max() - 1.0) > 1e-12:  # :648
                raise RuntimeError("The function `" + f.__name__ + "` generates values outside the expected range of zero to one.")
    return A, names
