"""Parity at BASELINE.json's FULL sizes, where the NumPy oracle cannot hold or finish the whole problem: the
device's results are checked on SAMPLES regenerated bit-exactly on the CPU (oracle/synth.py is the definition of
the data; csrc/generate.cu its device twin, test_generator_bit_exact).

A GRM entry K[i, i'] needs rows i and i' over ALL p markers and the column means: the sampled rows are regenerated
over the whole marker range, the means are the device's own, themselves checked against the CPU on a sample of
columns regenerated over all n rows -- so every quantity that enters the reference value is pinned to the CPU."""
import numpy as np
import pytest

from oracle import gwas_oracle as go, synth

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def sampled_rows(seed, n, p, kind, rows, chunk=50_000):
    R = np.empty((rows.size, p))
    for j0 in range(0, p, chunk):
        c = min(chunk, p - j0)
        R[:, j0:j0 + c] = synth.block(seed, n, j0, c, kind, rows=rows)
    return R


def check_means_on_sampled_columns(st, seed, n, p, kind, rng, count=192):
    cols = np.unique(np.concatenate([[0, 1, p - 1], rng.integers(0, p, count)]))
    A = np.asfortranarray(np.hstack([synth.block(seed, n, int(j), 1, kind) for j in cols]))
    mu, v = go.column_std(A)
    np.testing.assert_allclose(st["mean"][cols], mu, rtol=1e-14, atol=1e-16)
    assert np.array_equal(st["keep"][cols], v > go.EPS)
    ok = v > go.EPS
    np.testing.assert_allclose(st["sd"][cols][ok], v[ok], rtol=1e-12)
    return cols, A


@pytest.mark.parametrize("storage", ["float64", "codes"])
def test_config1_grm_full_size_sampled_rows(gbm, storage):
    """BASELINE configs[1]: grmsimple, diploid, n = 5,000 x p = 100,000 (centred X X' / p, north_star's definition of
    the call at /root/reference/src/gwas.jl:124) -- FP64 DMMA kernel and the exact INT8 tensor-core kernel."""
    n, p, seed, kind = 5000, 100_000, 42, synth.KIND_DIPLOID
    rng = np.random.default_rng(1)
    dm = gbm.DeviceMatrix.generate(seed, n, p, kind)
    if storage == "codes":
        f64, dm = dm, dm.pack()
        f64.free()
        assert dm is not None
    st = dm.colstats()
    check_means_on_sampled_columns(st, seed, n, p, kind, rng)
    K, tf = dm.grm(gbm._lib.GRM_SIMPLE, 2, 0)
    dm.free()
    assert np.array_equal(K, K.T)
    rows = np.unique(np.concatenate([[0, 1, 127, 128, n - 1], rng.integers(0, n, 27)]))
    Z = sampled_rows(seed, n, p, kind, rows) - st["mean"][None, :]
    want = (Z @ Z.T) / p
    got = K[np.ix_(rows, rows)]
    assert np.max(np.abs(got - want)) <= RTOL * np.abs(want).max()
    # entries far from the diagonal are differences of large terms: still 1e-9 of their own size where they are not ~0
    big = np.abs(want) > 1e-4 * np.abs(want).max()
    assert np.max(np.abs(got - want)[big] / np.abs(want)[big]) < RTOL


def test_config3_tetraploid_full_size(gbm):
    """BASELINE configs[3]: grmploidyaware + gwaslmm, tetraploid frequencies, n = 2,000 x p = 500,000 (the
    ploidy-aware branch /root/reference/src/gwas.jl:117-121), through the one-call multi-GPU entry point on a group of
    one GPU.  GRM on sampled rows, ploidy, filter count, and z on sampled markers against the oracle's closed form
    with the ORACLE's PC1 (LAPACK SVD of the standardised GRM)."""
    from gbm_b200 import multigpu

    n, p, seed, kind = 2000, 500_000, 42, synth.KIND_TETRAPLOID
    rng = np.random.default_rng(2)
    y = synth.phenotype(seed, n, p, kind)
    ys = (y - y.mean()) / y.std(ddof=1)
    grp = multigpu.Group.local(1)
    try:
        sm = multigpu.ShardedMatrix.generate(grp, seed, n, p, kind, pack=False)
        st = sm.colstats()
        cols, A = check_means_on_sampled_columns(st, seed, n, p, kind, rng)
        assert int(round(1.0 / st["min_nonzero_kept"])) == 4  # Int(round(1 / minimum(G[G .!= 0.0])))  gwas.jl:119
        res = sm.gwas(ys, model=gbm._lib.MODEL_LMM, grm_type=gbm._lib.GRM_PLOIDY_AWARE)
        assert res["timing"]["ploidy"] == 4
        assert np.array_equal(res["idx_cols"], st["idx_cols"]) and res["idx_cols"].size == st["keep"].sum()
        K, _ = sm.grm(gbm._lib.GRM_PLOIDY_AWARE, 4, 0)
        sm.free()
    finally:
        grp.free()
    rows = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, 22)]))
    q = st["mean"]
    Z = sampled_rows(seed, n, p, kind, rows) - q[None, :]
    want = 4.0 * (Z @ Z.T) / np.sum(q * (1.0 - q))
    got = K[np.ix_(rows, rows)]
    assert np.max(np.abs(got - want)) <= RTOL * np.abs(want).max()
    # z of `x` in y ~ 1 + PC1 + x + (1|entries) (gwas.jl:358-385; closed form SURVEY App. A.3) on the sampled markers
    pc = go.pca_pc1(go.standardise_K(K))
    pcg = res["pc1"] if res["pc1"] @ pc > 0 else -res["pc1"]
    assert np.max(np.abs(pcg - pc)) < 1e-9
    mu, v = go.column_std(A)
    ok = v > go.EPS
    ref = go.scan_closed_form(A[:, ok], ys, pc)
    z = res["stat"][cols[ok]]
    assert np.max(np.abs(z - ref["stat_lmm"]) / np.maximum(np.abs(ref["stat_lmm"]), 1e-4 * np.abs(ref["stat_lmm"]).max())) < RTOL
