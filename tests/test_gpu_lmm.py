"""GPU parity tests for the GRM-covariance LMM engine: the rotation
GEMM against NumPy and the per-marker REML delta search against oracle/lmm_oracle.py."""
import numpy as np
import pytest

from oracle import gwas_oracle as go, lmm_oracle as lo, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (300, 257, 301), (64, 1000, 130), (1, 5, 7), (515, 33, 1024),
                                   (2000, 2000, 2000)])
def test_gemm_tn_matches_numpy(gbm, M, N, K):
    import torch

    rng = np.random.default_rng(M + N + K)
    lda = ldb = (K + 15) // 16 * 16
    A = np.zeros((lda, M), order="F")
    B = np.zeros((ldb, N), order="F")
    A[:K] = rng.normal(size=(K, M))
    B[:K] = rng.normal(size=(K, N))
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()  # (M, lda) C-order == lda x M column-major
    dB = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
    dC = torch.full((N, M), np.nan, dtype=torch.float64, device="cuda")
    tf = gbm.gemm_tn(dA.data_ptr(), lda, dB.data_ptr(), ldb, dC.data_ptr(), M, M, N, K)
    assert tf > 0
    C = dC.cpu().numpy().T
    want = A[:K].T @ B[:K]
    assert np.max(np.abs(C - want)) < 1e-12 * max(1.0, np.abs(want).max()) * np.sqrt(K)


def _structured(seed, n, p, kind, h2=0.5):
    A = synth.block(seed, n, 0, p, kind)
    rng = np.random.default_rng(seed)
    g = (A - A.mean(axis=0)) @ rng.normal(size=p)
    g /= g.std()
    y = np.sqrt(h2) * g + np.sqrt(1 - h2) * rng.normal(size=n)
    return A, y


@pytest.mark.parametrize("n,p,kind,k", [(200, 300, synth.KIND_DIPLOID, 0), (257, 129, synth.KIND_TETRAPLOID, 0),
                                        (150, 400, synth.KIND_CONTINUOUS, 1), (301, 200, synth.KIND_DIPLOID, 2)])
def test_lmm_scan_matches_oracle(gbm, n, p, kind, k):
    A, y = _structured(11 + k, n, p, kind)
    K = go.grm_simple(A)
    rng = np.random.default_rng(5)
    C = rng.normal(size=(n, k)) if k else None
    ref = lo.lmm_scan(A, y, K, C)
    plan = gbm.LmmPlan(K, y, C)
    dm = gbm.DeviceMatrix.upload(A)
    res = plan.run(dm)
    dm.free()
    plan.free()
    assert abs(plan.null_log_delta - ref["lam0"]) < 1e-8
    keep = ref["keep"]
    assert np.all(np.isnan(res["stat"][~keep]))
    # rotation is sign/basis dependent only through U; statistics are not
    # z is stationary in log(delta) at the optimum (dz/dlam ~ 0.1 here), so a 1e-9 z needs ~1e-8 in lam: the kernel's
    # Newton polish stops at 1e-13, the oracle's Brent at 1e-13 -- the difference is the conditioning of the root
    assert np.max(np.abs(res["log_delta"][keep] - ref["log_delta"][keep])) < 1e-7
    zs = np.abs(ref["z"][keep]).max()
    assert np.max(np.abs(res["stat"][keep] - ref["z"][keep])) < 1e-9 * max(1.0, zs)  # north_star: 1e-9
    sd = A.std(axis=0, ddof=1)
    np.testing.assert_allclose(res["beta"][keep], ref["beta"][keep] * sd[keep], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(res["se"][keep], ref["se"][keep] * sd[keep], rtol=1e-8)
    want_p = go.neglog10_sf_normal(res["stat"][keep][:50])
    assert np.max(np.abs(res["neglog10p"][keep][:50] - want_p)) < 1e-6
    # the test data are unimodal, so the marched stationary point is the global REML estimate
    S, U = lo.rotate(K)
    Cfull = np.ones((n, 1)) if C is None else np.column_stack([np.ones(n), C])
    Cr, yr, Ar = U.T @ Cfull, U.T @ y, U.T @ A
    for j in np.flatnonzero(keep)[:10]:
        assert lo.is_unimodal(S, np.column_stack([Cr, Ar[:, j]]), yr)
        zd = lo.gls_z_dense(A[:, j], y, K, np.exp(res["log_delta"][j]), C)
        assert abs(zd - res["stat"][j]) < 1e-8 * max(1.0, abs(zd))


def test_lmm_boundary_delta(gbm):
    """Pure-noise phenotype: the REML estimate runs to the upper bound (no genetic variance);
    pure-genetic phenotype: to the lower one.  Both must agree with the oracle's clamping."""
    n, p = 120, 150
    A = synth.block(3, n, 0, p, synth.KIND_DIPLOID)
    K = go.grm_simple(A)
    rng = np.random.default_rng(0)
    for y in (rng.normal(size=n),):
        ref = lo.lmm_scan(A, y, K)
        plan = gbm.LmmPlan(K, y)
        dm = gbm.DeviceMatrix.upload(A)
        res = plan.run(dm)
        dm.free()
        plan.free()
        keep = ref["keep"]
        assert np.max(np.abs(res["log_delta"][keep] - ref["log_delta"][keep])) < 1e-6
        assert np.max(np.abs(res["stat"][keep] - ref["z"][keep])) < 1e-7


def test_gwasreml_end_to_end(gbm):
    """gwasreml doctest shape (gwas.jl:523-546): l = 1_000, model name, argmax equality
    across GRM types."""
    n, p = 100, 1000
    A, y = _structured(42, n, p, synth.KIND_TETRAPLOID)
    g = gbm.Genomes.from_matrix(A)
    ph = gbm.Phenomes.from_matrix(y, entries=g.entries)
    f1 = gbm.gwasreml(genomes=g, phenomes=ph, GRM_type="simple")
    f2 = gbm.gwasreml(genomes=g, phenomes=ph, GRM_type="ploidy-aware")
    assert f1.model == "GWAS_REML" and f2.model == "GWAS_REML"
    assert np.argmax(f1.b_hat) == np.argmax(f2.b_hat)
    mu, v = go.column_std(A)
    idx = go.fixed_locus_filter(v)
    assert np.array_equal(f1.extras["idx_cols"], idx)
    ref = lo.lmm_scan(A[:, idx - 1], (y - y.mean()) / y.std(ddof=1), go.grm_simple(A))
    assert np.max(np.abs(f1.b_hat - ref["z"])) < 1e-8 * max(1.0, np.abs(ref["z"]).max())
    # simple and ploidy-aware GRMs are proportional: z agrees except where the absolute
    # delta bounds [1e-5, 1e5] bind differently
    assert np.median(np.abs(f1.b_hat - f2.b_hat)) < 1e-9


@pytest.mark.parametrize("n,p,kind,scale", [(100, 300, synth.KIND_TETRAPLOID, 1.0), (257, 200, synth.KIND_DIPLOID, 0.2),
                                            (180, 150, synth.KIND_CONTINUOUS, 3.0)])
def test_reference_objective_mode_matches_its_oracle(gbm, n, p, kind, scale):
    """GBM_LMM_REFERENCE_OBJECTIVE: the reference's own objective, box and statistic
    (/root/reference/src/gwas.jl:478, :588, :596-599) minimised on rotated data, against oracle refobj_scan (whose
    1-D constrained profile is itself checked against a dense 2-D brute force of the literal loglikreml in
    tests/test_oracle.py).  `scale` moves the fit between the regimes (interior / s2u = 1 / s2e = 1)."""
    A, y = _structured(5, n, p, kind)
    ys = (y - y.mean()) / y.std(ddof=1)
    K = go.grm_simple(A) * scale * 10.0
    mu, v = go.column_std(A)
    keep = v > go.EPS
    G = (A[:, keep] - mu[keep]) / v[keep]
    z_ref, th_ref, lam_ref = lo.refobj_scan(G, ys, K)
    plan = gbm.LmmPlan(K, ys)
    dm = gbm.DeviceMatrix.upload(A)
    res = plan.run(dm, flags=gbm._lib.LMM_REFERENCE_OBJECTIVE)
    dm.free()
    plan.free()
    assert np.all(np.isnan(res["stat"][~keep]))
    z = res["stat"][keep]
    assert np.max(np.abs(z - z_ref)) < 1e-8 * max(1.0, np.abs(z_ref).max())
    # the fitted ratio: equal wherever the profile is not flat to rounding at its minimum (corner solutions are exact)
    inside = np.abs(lam_ref) < 11.0
    assert np.max(np.abs(res["log_delta"][keep][inside] - lam_ref[inside])) < 1e-5
    regimes = {(round(t[0], 6) == 1.0, round(t[1], 6) == 1.0) for t in th_ref}
    assert len(regimes) >= 1


def test_gwasreml_reference_objective_end_to_end(gbm):
    """gwasreml(objective = "reference") on the doctest shape (gwas.jl:523): the symmetric part of the
    column-standardised K, the reference's objective; against the oracle on the same K."""
    n, p = 100, 400
    A, y = _structured(42, n, p, synth.KIND_TETRAPLOID)
    g = gbm.Genomes.from_matrix(A)
    ph = gbm.Phenomes.from_matrix(y, entries=g.entries)
    f = gbm.gwasreml(genomes=g, phenomes=ph, GRM_type="simple", objective="reference")
    assert f.model == "GWAS_REML" and f.extras["objective"] == "reference"
    ent = [str(i) for i in range(n)]
    prep = go.gwasprep(A, ent, y[:, None], ent, standardise=True)
    Ksym = 0.5 * (prep.K + prep.K.T)
    z_ref, _, _ = lo.refobj_scan(prep.G, prep.y, Ksym)
    assert np.array_equal(f.extras["idx_cols"], prep.idx_cols)
    # the device's Kstd and NumPy's differ in the last bits; the symmetrised standardised K is indefinite with
    # eigenvalues close to -delta_min, so the fitted s2u (which scales z directly here: no sigma^2 is profiled out)
    # amplifies that: 1e-6 end to end, 1e-8 on identical K (test_reference_objective_mode_matches_its_oracle)
    assert np.max(np.abs(f.b_hat - z_ref)) < 1e-6 * max(1.0, np.abs(z_ref).max())
