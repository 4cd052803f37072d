"""GPU parity tests: the CUDA path (through the C ABI, via ctypes) against the CPU oracle on
the same seeded inputs.  Tolerances are BASELINE.json's: statistics / GRM within 1e-9
relative, -log10 p within 1e-6 absolute, the marker filter bit-exact."""
import numpy as np
import pytest

from oracle import cbind, gwas_oracle as go, synth

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def rel_err(a, b, floor=1e-4):
    """max |a-b| / max(|b|, floor*scale): relative, with a floor of 1e-4 of the largest value.  A statistic near 0
    is an exact cancellation of sums of n products: FP64 bounds its ABSOLUTE error (~ eps sqrt(n) scale ~ 1e-14 scale
    at n = 10,000, in the oracle as much as in the kernel), so "1e-9 relative" is only meaningful down to entries
    ~1e-5 of the largest; the floor makes the test demand 1e-13 of the scale in absolute terms below that."""
    a, b = np.asarray(a), np.asarray(b)
    scale = max(float(np.nanmax(np.abs(b))), 1e-300)
    return float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), floor * scale)))


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [synth.KIND_DIPLOID, synth.KIND_TETRAPLOID, synth.KIND_CONTINUOUS])
@pytest.mark.parametrize("n,p,col0", [(300, 257, 0), (1001, 64, 5), (2, 17, 0), (4099, 33, 1000)])
def test_generator_bit_exact(gbm, kind, n, p, col0):
    dm = gbm.DeviceMatrix.generate(42, n, p, kind, col0)
    got = dm.download()
    dm.free()
    want = synth.block(42, n, col0, p, kind)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,p", [(300, 1000), (257, 31), (2, 5), (513, 16), (1025, 200), (10000, 64)])
@pytest.mark.parametrize("kind", [synth.KIND_DIPLOID, synth.KIND_CONTINUOUS])
def test_upload_download_roundtrip_and_colstats(gbm, n, p, kind):
    A = synth.block(7, n, 0, p, kind)
    dm = gbm.DeviceMatrix.upload(A)
    assert np.array_equal(dm.download(), A)
    st = dm.colstats()
    dm.free()
    mu, v = go.column_std(A)
    idx = go.fixed_locus_filter(v)
    assert np.array_equal(st["idx_cols"], idx)  # bit-exact filter, 1-based ascending
    assert np.array_equal(st["keep"], v > go.EPS)
    np.testing.assert_allclose(st["mean"], mu, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(st["sd"][idx - 1], v[idx - 1], rtol=1e-12)
    assert np.all(st["sd"][~st["keep"]] <= go.EPS)
    # ploidy probe: minimum(G[G .!= 0]) over kept columns (gwas.jl:119)
    G = A[:, idx - 1]
    if G.size and (G != 0).any():
        assert st["min_nonzero_kept"] == G[G != 0].min()


def test_constant_non_dyadic_column_is_filtered(gbm):
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.random((777, 40)))
    A[:, 3] = 0.3
    A[:, 11] = 1.0 / 3.0
    A[:, 20] = 0.7
    dm = gbm.DeviceMatrix.upload(A)
    st = dm.colstats()
    dm.free()
    assert not st["keep"][[3, 11, 20]].any() and st["keep"].sum() == 37
    assert np.all(st["sd"][[3, 11, 20]] == 0.0)
    _, v = go.column_std(A)
    assert np.array_equal(st["idx_cols"], go.fixed_locus_filter(v))


def _problem(seed, n, p, kind):
    A = synth.block(seed, n, 0, p, kind)
    y = synth.phenotype(seed, n, p, kind)
    ys = (y - y.mean()) / y.std(ddof=1)
    rng = np.random.default_rng(seed)
    pc = rng.normal(size=n)
    pc -= pc.mean()
    pc /= np.linalg.norm(pc)
    return A, ys, pc


@pytest.mark.parametrize("n,p,kind", [
    (300, 2000, synth.KIND_CONTINUOUS),   # BASELINE config 1 shape (reduced p)
    (300, 10000, synth.KIND_TETRAPLOID),  # config 1 full size, doctest-style tetraploid rounding
    (1000, 500, synth.KIND_DIPLOID),
    (255, 33, synth.KIND_DIPLOID),        # ragged: n not a multiple of the 256-row stage, p not of 16
    (257, 16, synth.KIND_CONTINUOUS),
    (2001, 47, synth.KIND_TETRAPLOID),    # odd n: re-pitched upload
    (5, 3, synth.KIND_CONTINUOUS),
])
def test_scan_matches_oracle(gbm, n, p, kind):
    A, ys, pc = _problem(11, n, p, kind)
    dm = gbm.DeviceMatrix.upload(A)
    ols = dm.scan(ys, pc[:, None], model=0)
    lmm = dm.scan(ys, pc[:, None], model=1)
    dm.free()
    mu, v = go.column_std(A)
    keep = v > go.EPS
    assert np.array_equal(ols["keep"], keep)
    assert np.array_equal(lmm["keep"], keep)
    ref = go.scan_closed_form(A[:, keep], ys, pc)
    # the literal per-marker pinv route of the reference (gwas.jl:241-245)
    G = (A[:, keep] - mu[keep]) / v[keep]
    lit = go.gwasols_literal(G, ys, pc)
    if n > 3:
        assert rel_err(ols["stat"][keep, 0], lit) < 1e-8  # the literal route itself carries ~1e-11
        assert rel_err(ols["stat"][keep, 0], ref["stat_ols"]) < RTOL
        assert rel_err(ols["beta"][keep, 0], ref["beta"]) < RTOL
        assert rel_err(ols["se"][keep, 0], ref["se_ols"], floor=0) < RTOL
    if n > 4:
        assert rel_err(lmm["stat"][keep, 0], ref["stat_lmm"]) < RTOL
        assert rel_err(lmm["se"][keep, 0], ref["se_lmm"], floor=0) < RTOL
    assert np.all(np.isnan(ols["stat"][~keep, 0]))


def test_scan_matches_c_twin_literal(gbm):
    n, p = 500, 3000
    A, ys, pc = _problem(3, n, p, synth.KIND_DIPLOID)
    so, sl, keep = cbind.gwasols_raw(A, ys, pc)
    dm = gbm.DeviceMatrix.upload(A)
    ols = dm.scan(ys, pc[:, None], model=0, want=("stat",))
    lmm = dm.scan(ys, pc[:, None], model=1, want=("stat",))
    dm.free()
    assert np.array_equal(ols["keep"], keep)
    assert rel_err(ols["stat"][keep, 0], so[keep]) < 1e-8
    assert rel_err(lmm["stat"][keep, 0], sl[keep]) < 1e-8


def test_scan_pvalues(gbm):
    n, p = 400, 1500
    A, ys, pc = _problem(5, n, p, synth.KIND_DIPLOID)
    dm = gbm.DeviceMatrix.upload(A)
    ols = dm.scan(ys, pc[:, None], model=0)
    lmm = dm.scan(ys, pc[:, None], model=1)
    dm.free()
    keep = ols["keep"]
    sel = np.flatnonzero(keep)[:: max(1, keep.sum() // 200)]
    want_t = go.neglog10_sf_t(ols["stat"][sel, 0], n - 1)        # TDist(n-1), gwas.jl:252
    want_z = go.neglog10_sf_normal(lmm["stat"][sel, 0])          # Normal(),  gwas.jl:392
    assert np.max(np.abs(ols["neglog10p"][sel, 0] - want_t)) < 1e-6
    assert np.max(np.abs(lmm["neglog10p"][sel, 0] - want_z)) < 1e-6


@pytest.mark.parametrize("df", [3.0, 30.0, 299.0, 9999.0, 1e6])
def test_neglog10_sf_grid(gbm, df):
    t = np.concatenate([np.linspace(0, 10, 41), np.array([12.0, 20.0, 37.5, 40.0, 80.0, 200.0, -3.0])])
    got_t = gbm.neglog10_sf(t, "t", df)
    got_z = gbm.neglog10_sf(t, "normal")
    want_t = go.neglog10_sf_t(t, df)
    want_z = go.neglog10_sf_normal(t)
    assert np.max(np.abs(got_t - want_t)) < 1e-6
    assert np.max(np.abs(got_z - want_z)) < 1e-6
    # p itself within 1e-9 relative where it is representable
    ok = want_t < 300
    assert np.max(np.abs(10.0 ** (want_t[ok] - got_t[ok]) - 1.0)) < 1e-9


@pytest.mark.parametrize("T,k", [(1, 0), (3, 1), (5, 2), (13, 1), (20, 1)])
def test_scan_multi_trait_and_covariates(gbm, T, k):
    n, p = 600, 300
    A = synth.block(9, n, 0, p, synth.KIND_CONTINUOUS)
    rng = np.random.default_rng(T * 10 + k)
    Y = rng.normal(size=(n, T))
    C = rng.normal(size=(n, k)) if k else None
    dm = gbm.DeviceMatrix.upload(A)
    res = dm.scan(Y, C, model=1)
    dm.free()
    mu, v = go.column_std(A)
    keep = v > go.EPS
    Q = np.hstack([np.ones((n, 1))] + ([C] if k else []))
    Qo, _ = np.linalg.qr(Q)
    d = A[:, keep] - mu[keep]
    Md = d - Qo @ (Qo.T @ d)
    xMx = np.einsum("ij,ij->j", Md, Md)
    for t in range(T):
        My = Y[:, t] - Qo @ (Qo.T @ Y[:, t])
        s = (My @ Md) / np.sqrt(xMx)
        z = s / np.sqrt((My @ My - s * s) / (n - k - 2))
        assert rel_err(res["stat"][keep, t], z) < RTOL


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,p,kind", [(300, 1000, synth.KIND_CONTINUOUS), (129, 77, synth.KIND_DIPLOID),
                                      (640, 4099, synth.KIND_TETRAPLOID), (1000, 20000, synth.KIND_DIPLOID)])
def test_grm_matches_oracle(gbm, n, p, kind):
    A = synth.block(21, n, 0, p, kind)
    dm = gbm.DeviceMatrix.upload(A)
    Ks, tf = dm.grm(0, 2, 0)
    Ku, _ = dm.grm(0, 2, 1)
    Kp, _ = dm.grm(1, 4, 0)
    dm.free()
    for got, want in ((Ks, go.grm_simple(A)), (Ku, go.grm_simple(A, center=False)), (Kp, go.grm_ploidy_aware(A, 4))):
        assert np.array_equal(got, got.T)  # mirrored exactly
        scale = np.abs(want).max()
        assert np.max(np.abs(got - want)) < RTOL * scale
        big = np.abs(want) > 1e-3 * scale
        assert np.max(np.abs(got[big] / want[big] - 1.0)) < RTOL


@pytest.mark.parametrize("n", [64, 300, 515])
def test_kstd_pc1_matches_oracle(gbm, n):
    A = synth.block(33, n, 0, 2000, synth.KIND_CONTINUOUS)
    K = go.grm_simple(A)
    Ks, pc, _ = gbm.kstd_pc1(K)
    want_Ks = go.standardise_K(K)
    np.testing.assert_allclose(Ks, want_Ks, rtol=1e-10, atol=1e-11)
    want_pc = go.pca_pc1(want_Ks)
    sgn = np.sign(pc @ want_pc)
    assert abs(np.linalg.norm(pc) - 1) < 1e-12 and abs(pc.sum()) < 1e-10
    assert np.max(np.abs(sgn * pc - want_pc)) < 1e-9


@pytest.mark.parametrize("n", [2048, 1301])
def test_pc1_cooperative_reorthogonalisation_equals_the_five_kernel_route(gbm, n, monkeypatch):
    """The Lanczos step's Gram-Schmidt (twice), alpha, beta and the next basis vector run as ONE cooperative kernel
    (reorth_kernel, csrc/lanczos.cu); GBM_PC1_NO_COOP=1 keeps the five separate launches.  Same arithmetic up to the
    order of the fixed-order sums: the vectors agree far inside the 1e-9 they both owe the oracle, and each route
    repeats its own bits."""
    A = synth.block(21, n, 0, 2 * n, synth.KIND_DIPLOID)
    K = go.grm_simple(A)
    want = go.pca_pc1(go.standardise_K(K))
    monkeypatch.delenv("GBM_PC1_SOLVER", raising=False)
    monkeypatch.delenv("GBM_PC1_NO_COOP", raising=False)
    _, pc_coop, _ = gbm.kstd_pc1(K, want_kstd=False)
    assert gbm.last_timing()["launches"] > 20
    _, pc_coop2, _ = gbm.kstd_pc1(K, want_kstd=False)
    monkeypatch.setenv("GBM_PC1_NO_COOP", "1")
    _, pc_five, _ = gbm.kstd_pc1(K, want_kstd=False)
    assert np.array_equal(pc_coop, pc_coop2)
    assert np.max(np.abs(pc_coop - np.sign(pc_coop @ pc_five) * pc_five)) < 1e-10
    for pc in (pc_coop, pc_five):
        assert np.max(np.abs(np.sign(pc @ want) * pc - want)) < 1e-9


@pytest.mark.parametrize("n,kind", [(1024, synth.KIND_DIPLOID), (1500, synth.KIND_CONTINUOUS), (2600, synth.KIND_TETRAPLOID),
                                    (1301, synth.KIND_DIPLOID)])
def test_pc1_lanczos_matches_oracle_and_cusolver(gbm, n, kind, monkeypatch):
    """n >= 1024: PC1 comes from the Lanczos solver (csrc/lanczos.cu).  The spectrum of the standardised GRM
    is a near-degenerate bulk (relative gap of the top eigenvalue ~3e-3), the hard case for an iterative
    solver: the vector must still match the oracle's LAPACK SVD and cuSOLVER's syevdx to 1e-9, and the scan
    statistics computed with it must match to 1e-9 relative."""
    A = synth.block(12, n, 0, 3 * n, kind)
    K = go.grm_simple(A)
    want_pc = go.pca_pc1(go.standardise_K(K))
    monkeypatch.delenv("GBM_PC1_SOLVER", raising=False)
    _, pc, eig_ms = gbm.kstd_pc1(K, want_kstd=False)
    if n % 2 == 0:  # (odd n runs Lanczos too -- on the padded Z, two-pass step -- asserted in the cooperative-route test)
        assert gbm.last_timing()["launches"] > 20  # the iterative solver ran (its steps are counted as launches)
    monkeypatch.setenv("GBM_PC1_SOLVER", "cusolver")
    _, pc_cs, _ = gbm.kstd_pc1(K, want_kstd=False)
    # the variant used for n >= 12,000: Lanczos on Z itself (Z Z' never formed)
    monkeypatch.setenv("GBM_PC1_SOLVER", "lanczos-gram")
    _, pc_gram, _ = gbm.kstd_pc1(K, want_kstd=False)
    assert gbm.last_timing()["launches"] > 20
    assert np.max(np.abs(np.sign(pc_gram @ want_pc) * pc_gram - want_pc)) < 1e-9
    monkeypatch.delenv("GBM_PC1_SOLVER", raising=False)
    _, pc_again, _ = gbm.kstd_pc1(K, want_kstd=False)
    assert np.max(np.abs(pc - pc_again)) < 1e-12  # B = Z Z' itself carries last-bit noise (atomic tile slices)
    assert abs(np.linalg.norm(pc) - 1) < 1e-12 and abs(pc.sum()) < 1e-9
    for other in (want_pc, pc_cs):
        sgn = np.sign(pc @ other)
        assert np.max(np.abs(sgn * pc - other)) < 1e-9
    y = synth.phenotype(12, n, 3 * n, kind)
    ys = (y - y.mean()) / y.std(ddof=1)
    dm = gbm.DeviceMatrix.upload(A[:, :500])
    a = dm.scan(ys, pc[:, None], model=1)["stat"]
    b = dm.scan(ys, want_pc[:, None], model=1)["stat"]
    dm.free()
    assert rel_err(a, b) < RTOL


# ---------------------------------------------------------------------------------------
def _structs(gbm, n, p, kind, seed=42):
    A = synth.block(seed, n, 0, p, kind)
    y = synth.phenotype(seed, n, p, kind)
    g = gbm.Genomes.from_matrix(A)
    ph = gbm.Phenomes.from_matrix(y, entries=g.entries)
    return A, y, g, ph


@pytest.mark.parametrize("GRM_type", ["simple", "ploidy-aware"])
def test_gwasols_gwaslmm_end_to_end(gbm, GRM_type):
    """BASELINE config 1: n=300, l=10_000, 1 trait -- the whole reference pipeline."""
    n, p = 300, 10000
    A, y, g, ph = _structs(gbm, n, p, synth.KIND_TETRAPLOID)
    f1 = gbm.gwasols(genomes=g, phenomes=ph, GRM_type=GRM_type)
    f2 = gbm.gwaslmm(genomes=g, phenomes=ph, GRM_type=GRM_type)
    assert f1.model == "GWAS_OLS" and f2.model == "GWAS_LMM"  # doctests gwas.jl:194-200, :317-323
    b_ref, prep, pc = go.gwasols(A, g.entries, y[:, None], ph.entries, GRM_type=GRM_type)
    z_ref, _, _ = go.gwaslmm(A, g.entries, y[:, None], ph.entries, GRM_type=GRM_type)
    assert np.array_equal(f1.extras["idx_cols"], prep.idx_cols)
    assert f1.b_hat_labels == [g.loci_alleles[j - 1] for j in prep.idx_cols]
    if GRM_type == "ploidy-aware":
        assert f1.extras["ploidy"] == prep.ploidy == 4
    assert rel_err(f1.b_hat, b_ref) < 1e-8
    assert rel_err(f2.b_hat, z_ref) < RTOL
    assert f1.checkdims() and f2.checkdims() and f1.metrics == {"": 0.0}


def test_gwasols_argmax_equal_across_grm_types(gbm):
    """The reference's own doctest assertion (gwas.jl:202-203, :325-326)."""
    _, _, g, ph = _structs(gbm, 300, 4000, synth.KIND_TETRAPLOID, seed=5)
    a = gbm.gwasols(genomes=g, phenomes=ph, GRM_type="simple")
    b = gbm.gwasols(genomes=g, phenomes=ph, GRM_type="ploidy-aware")
    assert np.argmax(a.b_hat) == np.argmax(b.b_hat)


def test_gwasprep_doctest_invariants(gbm):
    """gwasprep doctest (gwas.jl:53-74)."""
    n, p = 300, 3000
    A, y, g, ph = _structs(gbm, n, p, synth.KIND_TETRAPLOID)
    G, ys, K, fit = gbm.gwasprep(genomes=g, phenomes=ph)
    assert np.all(np.abs(G.mean(axis=0)) < 1e-10)
    assert np.all(np.abs(G.std(axis=0, ddof=1) - 1) < 1e-10)
    assert abs(ys.mean()) < 1e-10 and abs(ys.std(ddof=1) - 1) < 1e-10
    assert G.shape[0] == ys.shape[0] and K.shape == (n, n)
    assert len(fit.entries) == n and fit.b_hat.shape[0] == G.shape[1]
    prep = go.gwasprep(A, g.entries, y[:, None], ph.entries)
    np.testing.assert_allclose(G, prep.G, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(K, prep.K, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(ys, prep.y, rtol=1e-13)


def test_index_subsets_and_errors(gbm):
    n, p = 200, 500
    A, y, g, ph = _structs(gbm, n, p, synth.KIND_CONTINUOUS)
    cols = np.arange(3, 400, 2, dtype=np.int64)
    f = gbm.gwasols(genomes=g, phenomes=ph, idx_loci_alleles=cols)
    b_ref, prep, _ = go.gwasols(A, g.entries, y[:, None], ph.entries, idx_loci_alleles=cols)
    assert np.array_equal(f.extras["idx_cols"], prep.idx_cols)
    assert rel_err(f.b_hat, b_ref) < 1e-8
    with pytest.raises(gbm.ArgumentError):
        gbm.gwasols(genomes=g, phenomes=ph, GRM_type="fancy")
    with pytest.raises(gbm.ArgumentError):
        gbm.gwasols(genomes=g, phenomes=ph, idx_loci_alleles=[0, 1])
    with pytest.raises(gbm.ArgumentError):
        gbm.gwasols(genomes=g, phenomes=ph, idx_entries=[1, n + 1])
    with pytest.raises(gbm.ArgumentError):  # dropped entries -> GRM/G row mismatch
        gbm.gwasols(genomes=g, phenomes=ph, idx_entries=list(range(1, 101)))


def test_scan_host_matches_resident_scan(gbm):
    n, p = 1000, 5000
    A, ys, pc = _problem(13, n, p, synth.KIND_DIPLOID)
    dm = gbm.DeviceMatrix.upload(A)
    a = dm.scan(ys, pc[:, None], model=1)
    dm.free()
    b = gbm.scan_host(A, ys, pc[:, None], model=1, pack=False)  # plain Float64 blocks: bit-identical
    for key in ("beta", "se", "stat", "neglog10p", "mean", "sd"):
        assert np.array_equal(a[key], b[key], equal_nan=True), key
    assert np.array_equal(a["keep"], b["keep"])
    c = gbm.scan_host(A, ys, pc[:, None], model=1)  # default: host-packed blocks, same results to rounding
    keep = a["keep"]
    assert np.array_equal(keep, c["keep"])
    for key in ("beta", "se", "stat"):
        assert np.nanmax(np.abs(a[key][keep] - c[key][keep])) < 1e-11 * np.nanmax(np.abs(a[key][keep])), key


def test_scan_is_deterministic_and_shard_invariant(gbm):
    """Per-marker results do not depend on how the columns are sharded."""
    n, p = 1500, 4000
    A, ys, pc = _problem(17, n, p, synth.KIND_DIPLOID)
    dm = gbm.DeviceMatrix.upload(A)
    full = dm.scan(ys, pc[:, None], model=0)
    again = dm.scan(ys, pc[:, None], model=0)
    dm.free()
    assert np.array_equal(full["stat"], again["stat"], equal_nan=True)
    parts = []
    for j0, j1 in ((0, 1000), (1000, 1777), (1777, 4000)):
        d = gbm.DeviceMatrix.upload(np.asfortranarray(A[:, j0:j1]))
        parts.append(d.scan(ys, pc[:, None], model=0)["stat"])
        d.free()
    assert np.array_equal(np.vstack(parts), full["stat"], equal_nan=True)


def test_grm_sharded_accumulate_equals_single(gbm):
    import torch

    n, p = 384, 6000
    A = synth.block(4, n, 0, p, synth.KIND_DIPLOID)
    dm = gbm.DeviceMatrix.upload(A)
    K1, _ = dm.grm(0, 2, 0)
    dm.free()
    dK = torch.zeros(n * n, dtype=torch.float64, device="cuda:0")
    for j0, j1 in ((0, 2500), (2500, 6000)):
        d = gbm.DeviceMatrix.upload(np.asfortranarray(A[:, j0:j1]))
        d.grm_accumulate(dK.data_ptr(), centre=True)
        d.free()
    gbm.grm_finalize(dK.data_ptr(), n, 1.0 / p)
    K2 = dK.cpu().numpy().reshape(n, n).T
    assert np.max(np.abs(K1 - K2)) < 1e-12 * np.abs(K1).max()


# ---------------------------------------------------------------------------------------
import glob
import os

GOLDEN = sorted(f for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(f).startswith("transform_"))  # those belong to test_*transform*.py


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_path_against_golden_fixtures(gbm, path):
    """tests/golden/*.npz (made by tests/golden/make_golden.py from the oracle)."""
    g = np.load(path)
    A, y = g["A"], g["y"]
    n, p = A.shape
    ge = gbm.Genomes.from_matrix(A)
    ph = gbm.Phenomes.from_matrix(y, entries=ge.entries)
    for grm_type, tag in (("simple", "s"), ("ploidy-aware", "p")):
        f1 = gbm.gwasols(genomes=ge, phenomes=ph, GRM_type=grm_type)
        f2 = gbm.gwaslmm(genomes=ge, phenomes=ph, GRM_type=grm_type)
        assert np.array_equal(f1.extras["idx_cols"], g[f"idx_cols_{tag}"])
        assert rel_err(f1.b_hat, g[f"b_ols_{tag}"]) < RTOL
        assert rel_err(f1.b_hat, g[f"b_ols_literal_{tag}"]) < 1e-8
        assert rel_err(f2.b_hat, g[f"z_lmm_{tag}"]) < RTOL
        assert rel_err(f1.extras["beta"], g[f"beta_{tag}"]) < RTOL
        assert rel_err(f1.extras["se"], g[f"se_ols_{tag}"], floor=0) < RTOL
        assert np.max(np.abs(f1.extras["neglog10p"] - g[f"nlp_t_{tag}"])) < 1e-6
        assert np.max(np.abs(f2.extras["neglog10p"] - g[f"nlp_z_{tag}"])) < 1e-6
        sgn = np.sign(f1.extras["pc1"] @ g[f"pc1_{tag}"])
        assert np.max(np.abs(sgn * f1.extras["pc1"] - g[f"pc1_{tag}"])) < 1e-9
    if "ploidy" in g:
        assert f1.extras["ploidy"] == int(g["ploidy"])
    G, ys, K, _ = gbm.gwasprep(genomes=ge, phenomes=ph, GRM_type="ploidy-aware")
    np.testing.assert_allclose(K, g["K_p"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(ys, g["ys"], rtol=1e-13)
    Ks = gbm.grmsimple(ge).genomic_relationship_matrix
    Ku = gbm.grmsimple(ge, centre=False).genomic_relationship_matrix
    Kp = gbm.grmploidyaware(ge, ploidy=4).genomic_relationship_matrix
    for got, want in ((Ks, g["grm_simple"]), (Ku, g["grm_simple_uncentred"]), (Kp, g["grm_ploidy4"])):
        assert np.max(np.abs(got - want)) < RTOL * np.abs(want).max()


def test_scan_plan_reuse_and_device_outputs(gbm):
    import torch

    n, p = 1200, 2500
    A, ys, pc = _problem(23, n, p, synth.KIND_DIPLOID)
    dm = gbm.DeviceMatrix.upload(A)
    ref = dm.scan(ys, pc[:, None], model=1)
    plan = gbm.ScanPlan(dm, ys, pc[:, None], model=1)
    stat_d = torch.full((p,), -7.0, dtype=torch.float64, device="cuda:0")
    nlp_h = np.empty(p)
    keep_d = torch.zeros(p, dtype=torch.uint8, device="cuda:0")
    for _ in range(3):
        tm = plan.run(stat=stat_d, neglog10p=nlp_h, keep=keep_d)
        assert tm["launches"] == 2
        assert np.array_equal(stat_d.cpu().numpy(), ref["stat"][:, 0], equal_nan=True)
        assert np.array_equal(nlp_h, ref["neglog10p"][:, 0], equal_nan=True)
        assert np.array_equal(keep_d.cpu().numpy().astype(bool), ref["keep"])
    plan.free()
    dm.free()


def test_large_shape_properties(gbm):
    """BASELINE-size rows (n = 10,000) on a column block: size-independent properties --
    the generator block regenerated on the CPU bit-exactly, sampled columns against the
    oracle, and linearity of the statistic's numerator in y."""
    n, p, col0 = 10000, 4096, 777_000
    dm = gbm.DeviceMatrix.generate(42, n, p, synth.KIND_DIPLOID, col0)
    sample = [0, 1, 17, 1000, 4095]
    for j in sample:
        assert np.array_equal(dm.download(j, 1)[:, 0], synth.block(42, n, col0 + j, 1, synth.KIND_DIPLOID)[:, 0])
    rng = np.random.default_rng(0)
    y1, y2 = rng.normal(size=n), rng.normal(size=n)
    pc = rng.normal(size=n)
    r1 = dm.scan(y1, pc[:, None], model=0)
    r2 = dm.scan(y2, pc[:, None], model=0)
    r12 = dm.scan(y1 + 2.0 * y2, pc[:, None], model=0)
    keep = r1["keep"]
    # stat_ols = x'My / sqrt(x'Mx) is linear in y
    lin = r1["stat"][keep, 0] + 2.0 * r2["stat"][keep, 0]
    assert np.max(np.abs(r12["stat"][keep, 0] - lin)) < 1e-9 * np.abs(lin).max()
    A = np.asfortranarray(np.hstack([synth.block(42, n, col0 + j, 1, synth.KIND_DIPLOID) for j in sample]))
    pcn = pc - pc.mean()
    ref = go.scan_closed_form(A, y1, pcn / np.linalg.norm(pcn))
    ok = np.array([keep[j] for j in sample])
    assert rel_err(r1["stat"][sample, 0][ok], ref["stat_ols"][ok]) < RTOL
    dm.free()


def test_c_abi_error_codes(gbm):
    """The C ABI reports argument errors as GBM_ERR_ARGUMENT (-> Julia ArgumentError) with a
    message, never crashes, and stays usable afterwards."""
    from ctypes import byref, c_double, c_int64, c_void_p

    from gbm_b200 import _lib

    lib = _lib.load()
    h = c_void_p()
    A = np.asfortranarray(np.random.default_rng(0).random((10, 4)))
    assert lib.gbm_matrix_upload(None, 10, 4, 10, byref(h)) == _lib.GBM_ERR_ARGUMENT
    assert lib.gbm_matrix_upload(_lib.ptr(A), 1, 4, 10, byref(h)) == _lib.GBM_ERR_ARGUMENT      # < 2 entries
    assert lib.gbm_matrix_upload(_lib.ptr(A), 10, 4, 5, byref(h)) == _lib.GBM_ERR_ARGUMENT      # lda < n
    assert b"leading dimension" in lib.gbm_last_error()
    assert lib.gbm_matrix_upload(_lib.ptr(A), 10, 4, 10, byref(h)) == _lib.GBM_OK
    K = np.empty((10, 10), order="F")
    tf = c_double()
    assert lib.gbm_grm(h, 7, 2, 0, _lib.ptr(K), byref(tf)) == _lib.GBM_ERR_ARGUMENT             # GRM_type (gwas.jl:101-107)
    assert b"GRM_type" in lib.gbm_last_error()
    assert lib.gbm_grm(h, 1, 0, 0, _lib.ptr(K), byref(tf)) == _lib.GBM_ERR_ARGUMENT             # ploidy
    y = np.ones(10)
    stat = np.empty(4)
    args = (None,) * 2 + (_lib.ptr(stat),) + (None,) * 4
    assert lib.gbm_scan(h, _lib.ptr(y), 1, 10, None, 0, 10, 0, 0, *args) == _lib.GBM_ERR_ARGUMENT  # no trait variance
    assert b"variance" in lib.gbm_last_error()
    assert lib.gbm_scan(h, _lib.ptr(y), 0, 10, None, 0, 10, 0, 0, *args) == _lib.GBM_ERR_ARGUMENT
    assert lib.gbm_scan(h, _lib.ptr(y), 1, 10, None, 0, 10, 9, 0, *args) == _lib.GBM_ERR_ARGUMENT  # model
    y = np.arange(10.0)
    assert lib.gbm_scan(h, _lib.ptr(y), 1, 10, None, 0, 10, 0, 0, *args) == _lib.GBM_OK
    assert np.all(np.isfinite(stat))
    idx = np.array([1, 2, 99], dtype=np.int64)
    h2 = c_void_p()
    assert lib.gbm_matrix_upload_indexed(_lib.ptr(A), 10, 4, 10, None, 10, _lib.ptr(idx), 3, byref(h2)) == _lib.GBM_ERR_ARGUMENT
    assert b"idx_loci_alleles" in lib.gbm_last_error()
    n, p, lda, d = c_int64(), c_int64(), c_int64(), c_void_p()
    assert lib.gbm_matrix_info(h, byref(n), byref(p), byref(lda), byref(d)) == 0 and (n.value, p.value) == (10, 4)
    assert lda.value % 16 == 0
    assert lib.gbm_matrix_free(h) == _lib.GBM_OK


def test_config4_tetraploid_pipeline_reduced(gbm):
    """BASELINE configs[3] shape (grmploidyaware + gwaslmm on tetraploid frequencies), n = 2,000,
    p reduced to what the oracle finishes in seconds; ploidy inference must return 4."""
    n, p = 2000, 3000
    A, y, g, ph = _structs(gbm, n, p, synth.KIND_TETRAPLOID, seed=4)
    f = gbm.gwaslmm(genomes=g, phenomes=ph, GRM_type="ploidy-aware")
    z_ref, prep, _ = go.gwaslmm(A, g.entries, y[:, None], ph.entries, GRM_type="ploidy-aware")
    assert f.extras["ploidy"] == 4 == prep.ploidy
    assert np.array_equal(f.extras["idx_cols"], prep.idx_cols)
    assert rel_err(f.b_hat, z_ref) < RTOL


def test_full_size_config3_sampled_parity(gbm):
    """BASELINE configs[2] at full size (n = 10,000 x p = 1,000,000, 80 GB in HBM): the scan's
    results on columns sampled across the whole range (including byte offsets beyond 2^32)
    against the oracle on the regenerated columns, the filter count, and the uncentred GRM on
    entries recomputed on the CPU from regenerated rows."""
    import torch

    n, p = 10000, 1_000_000
    free, _ = torch.cuda.mem_get_info(0)
    if free < 95e9:
        pytest.skip("needs ~90 GB of free HBM")
    dm = gbm.DeviceMatrix.generate(42, n, p, synth.KIND_DIPLOID)
    rng = np.random.default_rng(0)
    y, pc = rng.normal(size=n), rng.normal(size=n)
    res = dm.scan(y, pc[:, None], model=1, want=("stat", "beta"))
    keep = res["keep"]
    cols = np.unique(np.concatenate([[0, 1, 15, 16, p - 1, p - 2, 536870912 // 10000 + 1], rng.integers(0, p, 48)]))
    A = np.asfortranarray(np.hstack([synth.block(42, n, int(j), 1, synth.KIND_DIPLOID) for j in cols]))
    mu, v = go.column_std(A)
    assert np.array_equal(keep[cols], v > go.EPS)
    np.testing.assert_allclose(res["mean"][cols], mu, rtol=1e-13, atol=1e-15)
    pcn = pc - pc.mean()
    ok = v > go.EPS
    ref = go.scan_closed_form(A[:, ok], y - y.mean(), pcn / np.linalg.norm(pcn))
    assert rel_err(res["stat"][cols[ok], 0], ref["stat_lmm"]) < RTOL
    assert rel_err(res["beta"][cols[ok], 0], ref["beta"]) < RTOL
    st = dm.colstats()
    assert st["idx_cols"].size == keep.sum() and np.all(np.diff(st["idx_cols"]) > 0)
    assert np.array_equal(st["idx_cols"], np.flatnonzero(keep) + 1)
    # uncentred GRM (A A'/p): a few entries from regenerated rows
    dK = torch.empty(n * n, dtype=torch.float64, device="cuda:0")
    dm.grm(0, 2, 1, out=dK)
    rows = np.array([0, 1, 4999, 9999])
    R = np.zeros((rows.size, p))
    for j0 in range(0, p, 50000):
        R[:, j0:j0 + 50000] = synth.block(42, n, j0, 50000, synth.KIND_DIPLOID, rows=rows)
    want = (R @ R.T) / p
    K = dK.view(n, n)  # column-major n x n == K' as a C-order view; K is symmetric
    rt = torch.as_tensor(rows, device="cuda:0")
    got = K.index_select(0, rt).index_select(1, rt).cpu().numpy()
    assert np.max(np.abs(got - want)) < 1e-11 * np.abs(want).max()
    # centred GRM (the default, north_star's "centred X X'"): same rows, the device's column means (checked above on
    # the sampled columns)
    dm.grm(0, 2, 0, out=dK)
    Zc = R - res["mean"][None, :]
    want_c = (Zc @ Zc.T) / p
    got_c = K.index_select(0, rt).index_select(1, rt).cpu().numpy()
    assert np.max(np.abs(got_c - want_c)) < RTOL * np.abs(want_c).max()
    dm.free()


@pytest.mark.parametrize("n,p,T,k", [(601, 300, 2, 1), (1027, 129, 7, 0), (32, 128, 30, 1), (35, 257, 45, 2), (4099, 77, 20, 1)])
def test_multi_trait_tensor_pipe_kernel_edges(gbm, n, p, T, k, monkeypatch):
    """More than two side vectors run on the FP64 tensor pipe (csrc/scan_mt.cu): ragged n (not a multiple of
    the 32-row stage or the 4-row MMA step), ragged p (not a multiple of the 128-marker tile), 31 side vectors
    in one pass and more than that in two, constant markers (SS exactly 0, filtered), and agreement with the
    FMA kernels (GBM_SCAN_NO_DMMA) where those can run."""
    A = synth.block(19, n, 0, p, synth.KIND_CONTINUOUS)
    A[:, 5] = 0.25
    A[:, p - 1] = 1.0 / 3.0  # constant, not dyadic
    rng = np.random.default_rng(n + T)
    Y = rng.normal(size=(n, T)) + 3.0
    C = rng.normal(size=(n, k)) if k else None
    dm = gbm.DeviceMatrix.upload(A)
    res = dm.scan(Y, C, model=0)
    mu, v = go.column_std(A)
    keep = v > go.EPS
    assert not keep[5] and not keep[p - 1]
    assert np.array_equal(res["keep"], keep)
    assert res["sd"][5] == 0.0 and res["sd"][p - 1] == 0.0
    np.testing.assert_allclose(res["mean"], mu, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(res["sd"][keep], v[keep], rtol=1e-11)
    Q = np.hstack([np.ones((n, 1))] + ([C] if k else []))
    Qo, _ = np.linalg.qr(Q)
    d = A[:, keep] - mu[keep]
    Md = d - Qo @ (Qo.T @ d)
    xMx = np.einsum("ij,ij->j", Md, Md)
    for t in range(T):
        My = Y[:, t] - Qo @ (Qo.T @ Y[:, t])
        s = (My @ Md) / np.sqrt(xMx)
        assert rel_err(res["stat"][keep, t], s) < RTOL, t
    if k + T <= 14:
        monkeypatch.setenv("GBM_SCAN_NO_DMMA", "1")
        ref = dm.scan(Y, C, model=0)
        monkeypatch.delenv("GBM_SCAN_NO_DMMA")
        assert np.array_equal(ref["keep"], res["keep"])
        assert rel_err(res["stat"][keep], ref["stat"][keep]) < 1e-11
    pk = None
    dm.free()


def test_multi_trait_on_packed_codes(gbm):
    n, p, T = 900, 2100, 20
    A = synth.block(4, n, 0, p, synth.KIND_TETRAPLOID)
    rng = np.random.default_rng(2)
    Y = rng.normal(size=(n, T))
    pc = rng.normal(size=(n, 1))
    dm = gbm.DeviceMatrix.upload(A)
    pk = dm.pack()
    a = dm.scan(Y, pc, model=1)
    b = pk.scan(Y, pc, model=1)
    assert np.array_equal(a["keep"], b["keep"])
    keep = a["keep"]
    for key in ("beta", "se", "stat", "mean", "sd"):  # the code kernel sums in code units: same values, last-bit rounding
        x, y = a[key][keep], b[key][keep]
        assert np.nanmax(np.abs(x - y)) <= 1e-12 * max(1.0, np.nanmax(np.abs(x))), key
    assert np.nanmax(np.abs(a["neglog10p"][keep] - b["neglog10p"][keep])) < 1e-9
    # ragged shapes on codes: n off the 32-row stage, p off the 256-marker tile, constant markers
    n2, p2 = 1027, 301
    B = synth.block(6, n2, 0, p2, synth.KIND_DIPLOID)
    B[:, 7] = 0.5
    Y2 = rng.normal(size=(n2, 7))
    d2 = gbm.DeviceMatrix.upload(B)
    k2 = d2.pack()
    a2, b2 = d2.scan(Y2, None, model=0), k2.scan(Y2, None, model=0)
    assert np.array_equal(a2["keep"], b2["keep"]) and not b2["keep"][7] and b2["sd"][7] == 0.0
    kk = a2["keep"]
    assert np.nanmax(np.abs(a2["stat"][kk] - b2["stat"][kk])) <= 1e-12 * np.nanmax(np.abs(a2["stat"][kk]))
    d2.free()
    k2.free()
    dm.free()
    pk.free()


@pytest.mark.parametrize("n", [300, 1001])
def test_degenerate_markers_follow_the_truncated_pinv(gbm, n):
    """`Vinv = pinv(X' * X)` (/root/reference/src/gwas.jl:242) on markers that make X = [1, PC1, g] rank deficient or
    nearly so: a marker that is an affine function of the covariate, its exact duplicate, a duplicate of another
    marker, a column that varies in one entry by a thousand ulps, and a constant column.  Checker: the LITERAL pinv route
    of the oracle (3x3 SVD with Julia's rtol = 3 eps), not the closed form."""
    rng = np.random.default_rng(n)
    p = 64
    A = synth.block(2, n, 0, p, synth.KIND_CONTINUOUS)
    pc = rng.normal(size=n)
    pc -= pc.mean()
    pc /= np.linalg.norm(pc)
    y = rng.normal(size=n) + 2.0 * pc * np.sqrt(n)
    ys = (y - y.mean()) / y.std(ddof=1)
    A[:, 3] = 0.5 + 0.25 * pc / np.abs(pc).max()      # collinear with PC1 (plus intercept)
    A[:, 4] = A[:, 3]                                 # ... and its duplicate
    A[:, 5] = 0.75 - 0.125 * pc / np.abs(pc).max()    # collinear, opposite sign
    A[:, 9] = A[:, 8]                                 # duplicate of an ordinary marker: not degenerate
    A[:, 12] = 0.3
    A[0, 12] = 0.3 + 1024 * np.spacing(0.3)           # varies by 1024 ulps in one entry (sd ~ 2e-15 > eps)
    A[:, 13] = 0.3                                    # constant, non-dyadic
    dm = gbm.DeviceMatrix.upload(A)
    res = dm.scan(ys, pc[:, None], model=0)
    lmm = dm.scan(ys, pc[:, None], model=1)
    dm.free()
    mu, v = go.column_std(A)
    keep = v > go.EPS
    assert np.array_equal(res["keep"], keep) and not keep[13]
    G = (A[:, keep] - mu[keep]) / v[keep]
    want = np.full(p, np.nan)
    want[keep] = go.gwasols_literal(G, ys, pc)
    for j in (3, 4, 5):  # truncated pseudo-inverse: finite, +-PC1'y (SURVEY App. A.2 "degenerate")
        assert np.isfinite(want[j]) and abs(abs(want[j]) - abs(pc @ ys)) < 1e-9 * abs(pc @ ys)
        assert abs(res["stat"][j, 0] - want[j]) < 1e-9 * abs(want[j]), j
        assert np.isnan(lmm["stat"][j, 0])           # gwaslmm: the mirror writes the 0.0 of a failed fit
    assert res["stat"][3, 0] == res["stat"][4, 0] and np.sign(res["stat"][5, 0]) == -np.sign(res["stat"][3, 0])
    ordinary = keep.copy()
    ordinary[[3, 4, 5, 12]] = False
    assert rel_err(res["stat"][ordinary, 0], want[ordinary], floor=1e-4) < RTOL
    assert res["stat"][8, 0] == res["stat"][9, 0]
    # the few-ulp column: kept (sd > eps) by the two-pass std of the reference and by the kernel alike; its
    # standardised values are rounding noise in the reference (a - mean(a) loses every bit), so only finiteness is
    # comparable
    assert keep[12] and np.isfinite(res["stat"][12, 0])
    # beta and SE of the truncated solution: b_g = c'h / (1 + c'c), sqrt(Vinv_gg) = |c| / (1 + c'c), c = PC1'g
    c = pc @ G[:, list(np.flatnonzero(keep)).index(3)]
    assert abs(res["beta"][3, 0] - c * (pc @ ys) / (1 + c * c)) < 1e-9 * abs(res["beta"][3, 0])
    assert abs(res["se"][3, 0] - abs(c) / (1 + c * c)) < 1e-9 * res["se"][3, 0]
