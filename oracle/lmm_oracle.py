"""CPU oracle for the GRM-covariance LMM scan  --  TEST
INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see below).

What it restates
----------------
The reference's GRM-as-covariance GWAS is ``gwasreml`` / ``loglikreml``
(/root/reference/src/gwas.jl:450-483, :549-613): per marker, V = sigma2_u*GRM + sigma2_e*I
(:464-471), variance components re-estimated per marker (:584-590), then the GLS statistic
b[end]/sqrt(inv(X'V^-1X)[end]) (:591-599) with X = [1, g_j] (:586).  BASELINE.json's north_star
asks for this "per-marker solve on rotated data ... as batched per-SNP delta searches rather
than a P3D shortcut".

Deliberate, documented differences from the reference's code (so parity is by definition
"unpinned" for this row):
  * the reference passes the COLUMN-STANDARDISED, hence non-symmetric, K (gwas.jl:130) as
    a covariance and minimises `0.5*log det V + y'Py + log det(X'V^-1 X)` (gwas.jl:478) with
    L-BFGS at g_tol = 1e-4 from theta = [0.5, 0.5] (:578, :588-590): neither a likelihood nor
    reproducible to better than the optimiser's tolerance, and O(n^3) per evaluation.
  * this oracle (and the CUDA engine) use the standard REML of the same model on the
    SYMMETRIC GRM: K = U S U', rotate once, profile out beta and sigma2_g, and maximise over
    delta = sigma2_e/sigma2_g per marker:
        LL(delta) = -1/2 [ (n-q) log R + sum_i log(s_i + delta) + log det(X~' W X~) ],
        W = diag(1/(s_i+delta)),  R = y~'Wy~ - b'X~'Wy~ ,  z = b_x / sqrt(R/(n-q) [ (X~'WX~)^-1 ]_xx ).
    The per-marker maximiser is the stationary point of LL in log(delta) reached from the
    null-model estimate delta0 (bracket marched in steps of 0.5 in log delta, then a root
    finder), clamped to [1e-5, 1e5].  On unimodal likelihoods (all the test data; asserted)
    this is the global REML estimate.
The z statistic is invariant to centring/scaling of the marker column and of y, so raw allele
frequencies are used (the reference standardises both first, gwas.jl:128-129).
"""
from __future__ import annotations

import numpy as np

LOG_DELTA_MIN = float(np.log(1e-5))
LOG_DELTA_MAX = float(np.log(1e5))
MARCH_STEP = 0.5


def rotate(K: np.ndarray):
    """K = U diag(S) U' (symmetric eigendecomposition, ascending S)."""
    Ksym = 0.5 * (K + K.T)
    S, U = np.linalg.eigh(Ksym)
    return S, U


def _gram(Z: np.ndarray, w: np.ndarray) -> np.ndarray:
    return (Z * w[:, None]).T @ Z


def reml_terms(lam: float, S: np.ndarray, Xr: np.ndarray, yr: np.ndarray):
    """LL, dLL/dlam and the GLS pieces at lam = log(delta) for rotated design Xr (n x q), yr."""
    n, q = Xr.shape
    d = np.exp(lam)
    w = 1.0 / (S + d)
    Z = np.column_stack([Xr, yr])
    G1 = _gram(Z, w)
    G2 = _gram(Z, w * w)
    A = G1[:q, :q]
    g = G1[:q, q]
    Ainv = np.linalg.inv(A)
    b = Ainv @ g
    R = G1[q, q] - g @ b
    sign, logdetA = np.linalg.slogdet(A)
    LL = -0.5 * ((n - q) * np.log(R) + np.sum(np.log(S + d)) + logdetA)
    # derivatives in lam: dG1/dlam = -d * G2
    v = np.concatenate([-b, [1.0]])
    dR = -d * (v @ G2 @ v)
    dlogdetA = -d * np.trace(Ainv @ G2[:q, :q])
    dL = d * np.sum(w)
    dLL = -0.5 * ((n - q) * dR / R + dL + dlogdetA)
    return LL, dLL, b, R, Ainv


def _root_from(lam0: float, f) -> float:
    """Stationary point of LL reached from lam0: march in the ascent direction in steps of
    MARCH_STEP until dLL changes sign (or a bound is hit), then Brent on dLL."""
    from scipy.optimize import brentq

    lam0 = min(max(lam0, LOG_DELTA_MIN), LOG_DELTA_MAX)
    f0 = f(lam0)
    if f0 == 0.0:
        return lam0
    direction = 1.0 if f0 > 0 else -1.0
    a, fa = lam0, f0
    while True:
        b = a + direction * MARCH_STEP
        b = min(max(b, LOG_DELTA_MIN), LOG_DELTA_MAX)
        fb = f(b)
        if fa * fb <= 0.0:
            lo, hi = (a, b) if a < b else (b, a)
            return float(brentq(f, lo, hi, xtol=1e-13, rtol=8.9e-16, maxiter=200))
        if b == LOG_DELTA_MIN or b == LOG_DELTA_MAX:
            return b  # monotone up to the bound: boundary estimate
        a, fa = b, fb


def null_model(S, Cr, yr, grid: int = 101):
    """delta0: global maximiser of the null REML (fixed effects Cr only) on a log grid,
    refined to the stationary point in the best cell."""
    lams = np.linspace(LOG_DELTA_MIN, LOG_DELTA_MAX, grid)
    ll = np.array([reml_terms(l, S, Cr, yr)[0] for l in lams])
    g = int(np.argmax(ll))
    f = lambda l: reml_terms(l, S, Cr, yr)[1]
    if g == 0 or g == grid - 1:
        lam0 = _root_from(lams[g], f)
    else:
        from scipy.optimize import brentq

        lo, hi = lams[g - 1], lams[g + 1]
        lam0 = float(brentq(f, lo, hi, xtol=1e-13)) if f(lo) * f(hi) < 0 else _root_from(lams[g], f)
    return lam0


def lmm_scan(A: np.ndarray, y: np.ndarray, K: np.ndarray, C: np.ndarray | None = None):
    """Per-marker REML LMM scan.  A: n x p raw allele frequencies, y: n, K: n x n symmetric
    GRM, C: n x k extra fixed covariates (intercept always included).
    Returns dict(beta, se, z, log_delta, lam0, keep)."""
    n, p = A.shape
    S, U = rotate(K)
    ones = np.ones((n, 1))
    Cfull = ones if C is None else np.column_stack([ones, np.asarray(C, dtype=np.float64).reshape(n, -1)])
    Cr = U.T @ Cfull
    yr = U.T @ np.asarray(y, dtype=np.float64)
    lam0 = null_model(S, Cr, yr)
    Ar = U.T @ A
    q = Cr.shape[1] + 1
    out = {k: np.full(p, np.nan) for k in ("beta", "se", "z", "log_delta")}
    sd = A.std(axis=0, ddof=1)
    keep = sd > np.finfo(np.float64).eps
    for j in range(p):
        if not keep[j]:
            continue
        Xr = np.column_stack([Cr, Ar[:, j]])
        f = lambda l: reml_terms(l, S, Xr, yr)[1]
        lam = _root_from(lam0, f)
        LL, dLL, b, R, Ainv = reml_terms(lam, S, Xr, yr)
        sg2 = R / (n - q)
        se = np.sqrt(sg2 * Ainv[-1, -1])
        out["beta"][j] = b[-1]
        out["se"][j] = se
        out["z"][j] = b[-1] / se
        out["log_delta"][j] = lam
    out["lam0"] = lam0
    out["keep"] = keep
    out["S"] = S
    return out


def gls_z_dense(a: np.ndarray, y: np.ndarray, K: np.ndarray, delta: float, C: np.ndarray | None = None):
    """Independent check of the rotation algebra: the same z from the un-rotated dense GLS
    (V = K + delta I, REML sigma2_g), O(n^3)."""
    n = a.shape[0]
    ones = np.ones((n, 1))
    X = np.column_stack([ones] + ([] if C is None else [np.asarray(C).reshape(n, -1)]) + [a])
    V = 0.5 * (K + K.T) + delta * np.eye(n)
    Vi = np.linalg.inv(V)
    XtViX = X.T @ Vi @ X
    b = np.linalg.solve(XtViX, X.T @ Vi @ y)
    r = y - X @ b
    sg2 = float(r @ Vi @ r) / (n - X.shape[1])
    return b[-1] / np.sqrt(sg2 * np.linalg.inv(XtViX)[-1, -1])


def is_unimodal(S, Xr, yr, grid: int = 201) -> bool:
    lams = np.linspace(LOG_DELTA_MIN, LOG_DELTA_MAX, grid)
    d = np.array([reml_terms(l, S, Xr, yr)[1] for l in lams])
    sign_changes = np.sum(np.sign(d[1:]) != np.sign(d[:-1]))
    return sign_changes <= 1
