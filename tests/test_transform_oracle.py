"""CPU suite for the transformation-screen oracle (oracle/transform_oracle.py): what the reference's
own doctests pin (transformation.jl:113-126, :298-316, :522-536), the closed form of the 2-column
least-squares problem, the rank-deficient branch of Julia's `\\`, and the committed golden fixtures."""
import os

import numpy as np
import pytest

from oracle import synth, transform_oracle as to

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _closed_form(z, y):
    zc = z - z.mean()
    return float(zc @ (y - y.mean()) / (zc @ zc))


def test_ols_slope_is_the_simple_regression_slope():
    rng = np.random.default_rng(0)
    for n in (5, 40, 301):
        z, y = rng.random(n), rng.normal(size=n) + 3.0
        assert abs(to.ols_slope(z, y) - _closed_form(z, y)) < 1e-12 * max(1.0, abs(_closed_form(z, y)))


def test_rank_deficient_feature_gets_the_minimum_norm_solution():
    """addnorm of complementary alleles is exactly constant: [1 z] has rank 1 and `\\` returns the
    minimum-norm solution, b2 = c ybar / (1 + c^2) -- not 0 and not NaN."""
    x = np.array([0.0, 0.5, 1.0, 0.5, 0.0, 1.0, 0.5]) + to.EPS
    z = to.addnorm(x, (1.0 - (x - to.EPS)) + to.EPS)
    assert np.all(z == z[0])
    y = np.array([3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.5])
    c = z[0]
    assert abs(to.ols_slope(z, y) - c * y.mean() / (1.0 + c * c)) < 1e-12


def test_doctest_pins_selected_feature_equals_f_of_its_locus():
    """transformation.jl:113-126: sqrt.(T[:, 1]) == af[:, idx] for f = x -> x^2 (1e-10)."""
    A = synth.block(3, 60, 0, 30, synth.KIND_CONTINUOUS)
    y = synth.phenotype(3, 60, 30, synth.KIND_CONTINUOUS, n_causal=4)
    beta, idx, T = to.transform1(to.square, A, y, n_new=30)
    assert T.shape[1] == idx.size > 0
    assert np.mean(np.sqrt(T[:, 0]) - A[:, idx[0] - 1]) < 1e-10
    assert np.all(np.abs(beta[idx - 1][:-1]) >= np.abs(beta[idx - 1][1:]))  # sortperm order, rev = true
    beta2, cnt, pairs, T2 = to.transform2(to.mult, A, y, n_new=25)
    assert np.all(np.diff(cnt) > 0)  # sort!(idx)
    i, j = pairs[0]
    np.testing.assert_allclose(T2[:, 0], (A[:, i - 1] + to.EPS) * (A[:, j - 1] + to.EPS), rtol=0, atol=1e-15)


def test_selection_requests_more_than_available_is_a_bounds_error():
    A = synth.block(3, 20, 0, 5, synth.KIND_CONTINUOUS)
    y = np.arange(20.0)
    with pytest.raises(IndexError):
        to.transform1(to.square, A, y, n_new=6)


def test_epistasisfeatures_ranges_and_names():
    """transformation.jl:522-536: new features appended, values in [0, 1], names f(locus...)."""
    A = synth.block(9, 50, 0, 16, synth.KIND_TETRAPLOID)
    y = synth.phenotype(9, 50, 16, synth.KIND_TETRAPLOID, n_causal=3)
    names = [f"chr1\t{j}\tA" for j in range(16)]
    B, new_names = to.epistasisfeatures(A, y, names, n_new=5, n_reps=2)
    assert B.shape[1] == len(new_names) > 16 and len(set(new_names)) == len(new_names)
    assert B.min() >= 0.0 and abs(B.max() - 1.0) <= 1e-12
    assert new_names[16].startswith("square(")


@pytest.mark.parametrize("name", ["transform_tetraploid_n40_l12", "transform_continuous_n57_l9"])
def test_oracle_reproduces_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    A, y = g["A"], g["y"]
    l = A.shape[1]
    for f in to.TRANSFORMATIONS1:
        beta, idx, T = to.transform1(f, A, y, n_new=min(6, l))
        np.testing.assert_allclose(beta, g[f"beta1_{f.__name__}"], rtol=1e-11, atol=1e-13)
        assert np.array_equal(idx, g[f"idx1_{f.__name__}"])
        np.testing.assert_allclose(T, g[f"T1_{f.__name__}"], rtol=1e-14, atol=0)
    for f in to.TRANSFORMATIONS2:
        beta, idx, pairs, T = to.transform2(f, A, y, n_new=10, commutative=True)
        np.testing.assert_allclose(beta, g[f"beta2_{f.__name__}_1"], rtol=1e-11, atol=1e-13)
        assert np.array_equal(idx, g[f"idx2_{f.__name__}_1"])
    # the fixture exercises the special branches
    assert g["beta1_square"][5] == 0.0  # low-variance locus skipped
    b = g["beta2_addnorm_0"].reshape(l, l)
    x = A[:, 2] + to.EPS
    c = ((x + (A[:, 3] + to.EPS)) / 2.0)[0]
    if "tetraploid" in name:  # complementary dosage alleles: exactly constant feature, minimum-norm solution
        assert np.all((x + (A[:, 3] + to.EPS)) / 2.0 == c)
        assert abs(b[2, 3] - c * y.mean() / (1 + c * c)) < 1e-10
    # closed form everywhere else
    X = A + to.EPS
    v = X.var(axis=0, ddof=1)
    bm = g["beta2_mult_0"].reshape(l, l)
    for i in range(l):
        for j in range(l):
            if v[i] >= 0.01 and v[j] >= 0.01:
                z = X[:, i] * X[:, j]
                if z.var() > 1e-12:
                    assert abs(bm[i, j] - _closed_form(z, y)) < 1e-9 * max(1.0, abs(bm[i, j]))
            else:
                assert bm[i, j] == 0.0
