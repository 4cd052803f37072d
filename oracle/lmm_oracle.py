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


# ======================================================================================
# LITERAL restatement of the reference's own loglikreml / gwasreml
# (/root/reference/src/gwas.jl:450-483, :549-613), line by line, dense O(n^3) per evaluation.
# Only usable at the doctest scale (n = 100, l = 1,000: gwas.jl:523).  It exists to MEASURE how
# far the engine's model is from what the reference's code computes (tests/test_oracle.py).
# ======================================================================================
_EPS = float(np.finfo(np.float64).eps)


def _julia_log(x: float) -> float:
    """Julia's log on a Float64: DomainError for x < 0 (caught by the caller -> Inf), -Inf at 0."""
    if np.isnan(x):
        return float("nan")
    if x < 0.0:
        raise ValueError("DomainError")
    if x == 0.0:
        return float("-inf")
    return float(np.log(x))


def _julia_pinv(M: np.ndarray) -> np.ndarray:
    """LinearAlgebra.pinv(M) with its default rtol = eps * min(size(M))."""
    return np.linalg.pinv(M, rcond=_EPS * min(M.shape))


def loglikreml_literal(theta, y: np.ndarray, X: np.ndarray, GRM: np.ndarray) -> float:
    """loglikreml (gwas.jl:450-483).  theta = [sigma2_e, sigma2_u]; GRM is whatever gwasprep returned
    (column-standardised, hence NON-symmetric, gwas.jl:130)."""
    n = y.shape[0]
    s2e, s2u = float(theta[0]), float(theta[1])  # :462, :465
    V = s2u * GRM + s2e * np.eye(n)  # :463-470
    V_inv = _julia_pinv(V)  # :471
    XtVi = X.T @ V_inv
    P = V_inv - (V_inv @ X @ np.linalg.inv(XtVi @ X) @ XtVi)  # :473
    y_reml = P @ y  # :474
    try:  # :476-480
        return 0.5 * _julia_log(float(np.linalg.det(V))) + float(y @ y_reml) + _julia_log(float(np.linalg.det(XtVi @ X)))
    except (ValueError, np.linalg.LinAlgError):
        return float("inf")


def gwasreml_statistic_literal(theta, y, X, GRM) -> float:
    """The statistic of gwasreml at the fitted theta (gwas.jl:591-599):
    b = pinv(X'V^-1X) X'V^-1 y, sigma2_b = inv(X'V^-1X), b[end] / sqrt(sigma2_b[end])."""
    n = y.shape[0]
    V = float(theta[1]) * GRM + float(theta[0]) * np.eye(n)
    V_inv = _julia_pinv(V)
    XtViX = X.T @ V_inv @ X
    b = _julia_pinv(XtViX) @ (X.T @ V_inv @ y)
    s2b = np.linalg.inv(XtViX)
    return float(b[-1] / np.sqrt(s2b[-1, -1]))


def gwasreml_literal(G: np.ndarray, y: np.ndarray, GRM: np.ndarray, markers=None):
    """gwasreml's marker loop (gwas.jl:577-599): X = [1, G[:, j]] (NO PC1, :586), minimise loglikreml over
    [eps, 1]^2 from theta = [0.5, 0.5] (:578, :588) with L-BFGS-B at g_tol = 1e-4 (:590; Optimization.LBFGS is
    the L-BFGS-B code, the same algorithm as SciPy's; the reference's gradient is Zygote's exact one, here
    SciPy's finite differences), then the statistic.  G, y, GRM as returned by gwasprep(standardise = true).
    Returns (z, theta) over `markers` (default: all)."""
    from scipy.optimize import minimize

    n, l = G.shape
    markers = range(l) if markers is None else markers
    ones = np.ones(n)
    z, th = [], []
    for j in markers:
        X = np.column_stack([ones, G[:, j]])
        f = lambda t: loglikreml_literal(t, y, X, GRM)
        sol = minimize(f, x0=np.array([0.5, 0.5]), method="L-BFGS-B", bounds=[(_EPS, 1.0), (_EPS, 1.0)],
                       options={"gtol": 1e-4, "ftol": 0.0, "maxiter": 200})
        th.append(sol.x.copy())
        z.append(gwasreml_statistic_literal(sol.x, y, X, GRM))
    return np.array(z), np.array(th)


# ======================================================================================
# The reference's OBJECTIVE, BOUNDS and STATISTIC on a SYMMETRIC K, through the eigen-rotation
# (the engine's "reference-objective" mode, GBM_LMM_REFERENCE_OBJECTIVE).  With K = U S U',
# V = t2 K + t1 I = U diag(t2 s_i + t1) U', so with c = t2, delta = t1 / t2, w_i = 1 / (s_i + delta):
#   0.5 log det V            = 0.5 n log c + 0.5 sum_i log(s_i + delta)
#   y'Py                     = Q(delta) / c,   Q = y~'Wy~ - g'A^-1 g,  A = X~'WX~,  g = X~'Wy~
#   log det(X'V^-1X)         = log det A(delta) - q log c
# For a fixed delta the objective is unimodal in c with its minimum at c* = Q / (n/2 - q); the box
# [eps, 1]^2 is c in [max(eps, eps/delta), min(1, 1/delta)], so the constrained minimiser in c is
# c* clamped to that interval and the problem is ONE-dimensional in lambda = log(delta).
# ======================================================================================
def refobj_terms(lam: float, S: np.ndarray, Xr: np.ndarray, yr: np.ndarray):
    n, q = Xr.shape
    d = float(np.exp(lam))
    w = 1.0 / (S + d)
    A = _gram(Xr, w)
    g = Xr.T @ (w * yr)
    Ainv = np.linalg.inv(A)
    b = Ainv @ g
    Q = float(yr @ (w * yr) - g @ b)
    sign, logdetA = np.linalg.slogdet(A)
    cstar = Q / (0.5 * n - q)
    lo, hi = max(_EPS, _EPS / d), min(1.0, 1.0 / d)
    c = min(max(cstar, lo), hi)
    f = (0.5 * n - q) * np.log(c) + 0.5 * float(np.sum(np.log(S + d))) + Q / c + logdetA
    return f, c, b, Ainv


def refobj_fit(S, Xr, yr, lam_lo=None, lam_hi=None):
    """argmin over lambda = log(delta) of the box-constrained profile (bounded Brent after a coarse grid: the
    profile is piecewise smooth).  Returns (theta = [t1, t2], z) with z = b_x / sqrt([A^-1]_xx * c) -- the
    reference's b[end] / sqrt(inv(X'V^-1X)[end]) (no sigma^2 factor, gwas.jl:596-599)."""
    from scipy.optimize import minimize_scalar

    # the engine's search interval: delta in [1e-5, 1e5] (theta_1 = eps makes V numerically singular), and
    # s_i + delta > 0 for an indefinite K (the symmetric part of the column-standardised GRM)
    if lam_lo is None:
        lam_lo = LOG_DELTA_MIN
        if S[0] < 0.0:
            lam_lo = max(lam_lo, float(np.log(-S[0] * (1.0 + 1e-6) + 1e-300)) + 1e-3)
    lam_hi = LOG_DELTA_MAX if lam_hi is None else lam_hi
    grid = np.linspace(lam_lo, lam_hi, 145)
    vals = np.array([refobj_terms(l, S, Xr, yr)[0] for l in grid])
    k = int(np.argmin(vals))
    a, b_ = grid[max(k - 1, 0)], grid[min(k + 1, grid.size - 1)]
    sol = minimize_scalar(lambda l: refobj_terms(l, S, Xr, yr)[0], bounds=(a, b_), method="bounded",
                          options={"xatol": 1e-12})
    lam = float(sol.x)
    if vals[k] < sol.fun:
        lam = float(grid[k])
    f, c, b, Ainv = refobj_terms(lam, S, Xr, yr)
    d = float(np.exp(lam))
    z = b[-1] / np.sqrt(Ainv[-1, -1] * c)
    return np.array([c * d, c]), float(z), lam, f


def refobj_scan(G: np.ndarray, y: np.ndarray, Ksym: np.ndarray, markers=None):
    """The reference-objective scan on the symmetric GRM: X = [1, g_j], per-marker fit of [t1, t2] in [eps, 1]^2."""
    n, l = G.shape
    S, U = rotate(Ksym)
    yr = U.T @ y
    one_r = U.T @ np.ones(n)
    Gr = U.T @ G
    markers = range(l) if markers is None else markers
    z, th, lams = [], [], []
    for j in markers:
        Xr = np.column_stack([one_r, Gr[:, j]])
        theta, zj, lam, _ = refobj_fit(S, Xr, yr)
        z.append(zj)
        th.append(theta)
        lams.append(lam)
    return np.array(z), np.array(th), np.array(lams)
