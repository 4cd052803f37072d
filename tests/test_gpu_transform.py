"""GPU parity tests for the transformation screens: csrc/transform.cu through
the C ABI against oracle/transform_oracle.py and the committed golden fixtures.  Effects within 1e-9
relative; the selection (indices, order) and the materialised features of the exactly-rounded
functions bit-exact."""
import os

import numpy as np
import pytest

from oracle import synth, transform_oracle as to

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-9


def _pairs(gbm):
    from gbm_b200 import transform as tr

    return tr, [(tr.square, to.square), (tr.invoneplus, to.invoneplus), (tr.log10epsdivlog10eps, to.log10epsdivlog10eps)], \
        [(tr.mult, to.mult), (tr.addnorm, to.addnorm), (tr.raise_, to.raise_)]


def _close(got, want, rtol=RTOL):
    scale = max(float(np.max(np.abs(want))), 1e-300)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * scale)
    assert float(np.max(err)) < rtol, float(np.max(err))


@pytest.mark.parametrize("name", ["transform_tetraploid_n40_l12", "transform_continuous_n57_l9"])
def test_golden_fixtures(gbm, name):
    tr, f1s, f2s = _pairs(gbm)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    A, y = g["A"], g["y"]
    l = A.shape[1]
    dm = gbm.DeviceMatrix.upload(A)
    for fd, fo in f1s:
        beta, idx = tr.transform1_screen(dm, y, fd, min(6, l))
        _close(beta, g[f"beta1_{fo.__name__}"])
        assert np.array_equal(idx, g[f"idx1_{fo.__name__}"])
        T = tr.transform1_apply(dm, fd, idx)
        np.testing.assert_allclose(T, g[f"T1_{fo.__name__}"], rtol=4e-16, atol=0)
    for fd, fo in f2s:
        for comm in (False, True):
            beta, cnt, vals = tr.transform2_screen(dm, y, fd, 10, commutative=comm, want_beta=True)
            want = g[f"beta2_{fo.__name__}_{int(comm)}"]
            _close(beta, want)
            assert np.array_equal(cnt, g[f"idx2_{fo.__name__}_{int(comm)}"])
            _close(vals, want[cnt - 1])
            T = tr.transform2_apply(dm, fd, cnt)
            np.testing.assert_allclose(T, g[f"T2_{fo.__name__}_{int(comm)}"], rtol=4e-16, atol=0)
    dm.free()


@pytest.mark.parametrize("n,l,kind", [(33, 70, synth.KIND_CONTINUOUS), (257, 129, synth.KIND_TETRAPLOID),
                                      (64, 64, synth.KIND_DIPLOID), (2, 5, synth.KIND_CONTINUOUS),
                                      (1001, 65, synth.KIND_CONTINUOUS)])
def test_ragged_shapes_against_the_oracle(gbm, n, l, kind):
    """n off the 32-row stage and odd, l off the 64-locus tile; use_abs, a non-default eps and
    variance threshold."""
    tr, f1s, f2s = _pairs(gbm)
    A = synth.block(21, n, 0, l, kind)
    rng = np.random.default_rng(n + l)
    y = 5.0 + A[:, 0] * 2.0 - A[:, l // 2] + rng.normal(size=n)
    dm = gbm.DeviceMatrix.upload(A)
    kw = dict(eps=1e-9, use_abs=True, var_threshold=0.02)
    for fd, fo in f1s:
        want, idx_w, T_w = to.transform1(fo, A, y, n_new=l, **kw)
        beta, idx = tr.transform1_screen(dm, y, fd, l, **kw)
        _close(beta, want)
        assert np.array_equal(beta == 0.0, want == 0.0)  # the same loci skipped
        if n > 2:
            assert np.array_equal(idx, idx_w)
    for fd, fo in f2s:
        want, idx_w, pairs, T_w = to.transform2(fo, A, y, n_new=min(l * l, 40), commutative=(fo is not to.raise_), **kw)
        beta, cnt, vals = tr.transform2_screen(dm, y, fd, min(l * l, 40), commutative=(fo is not to.raise_), want_beta=True, **kw)
        _close(beta, want)
        assert np.array_equal(beta == 0.0, want == 0.0)
        if n > 2:
            assert np.array_equal(cnt, idx_w)
            T = tr.transform2_apply(dm, fd, cnt, eps=1e-9, use_abs=True)
            np.testing.assert_allclose(T, T_w, rtol=4e-16, atol=0)
    dm.free()


def test_exactly_rounded_features_are_bit_exact_and_cleaned(gbm):
    """square / mult / addnorm / invoneplus are single correctly-rounded operations: T must equal the
    oracle's bit for bit, including the eps clean-up to exact 0 and 1 (transformation.jl:223-227)."""
    tr, _, _ = _pairs(gbm)
    A = synth.block(2, 90, 0, 20, synth.KIND_DIPLOID)  # values 0, 0.5, 1: f(x + eps) lands within eps of 0 and 1
    dm = gbm.DeviceMatrix.upload(A)
    idx = np.arange(1, 21, dtype=np.int64)
    X = A + to.EPS
    for fd, fo in ((tr.square, to.square), (tr.invoneplus, to.invoneplus)):
        T = tr.transform1_apply(dm, fd, idx)
        assert np.array_equal(T, to._clean(fo(X), to.EPS))
    assert (tr.transform1_apply(dm, tr.square, idx) == 0.0).any()  # (0 + eps)^2 < eps: cleaned to exact 0
    cnt = np.array([1, 2, 21, 47, 400], dtype=np.int64)
    for fd, fo in ((tr.mult, to.mult), (tr.addnorm, to.addnorm)):
        T = tr.transform2_apply(dm, fd, cnt)
        want = np.stack([fo(X[:, (c - 1) // 20], X[:, (c - 1) % 20]) for c in cnt], axis=1)
        assert np.array_equal(T, to._clean(want, to.EPS))
    dm.free()


def test_pairwise_screen_at_size_properties(gbm):
    """l = 1500, n = 2000 (2.25 M regressions): symmetry of commutative functions is bitwise, the
    non-commutative upper triangle equals the commutative run, sampled pairs match the closed form."""
    tr, _, _ = _pairs(gbm)
    n, l = 2000, 1500
    dm = gbm.DeviceMatrix.generate(5, n, l, synth.KIND_TETRAPLOID)
    A = dm.download()
    rng = np.random.default_rng(1)
    y = 3.0 + A[:, 10] * A[:, 700] * 4.0 + A[:, 3] + rng.normal(size=n)
    full, cnt, vals = tr.transform2_screen(dm, y, tr.mult, 1000, want_beta=True)
    comm, cnt_c, _ = tr.transform2_screen(dm, y, tr.mult, 1000, commutative=True, want_beta=True)
    B, C = full.reshape(l, l), comm.reshape(l, l)
    assert np.array_equal(B, B.T)
    iu = np.triu_indices(l)
    assert np.array_equal(B[iu], C[iu]) and not C[np.tril_indices(l, -1)].any()
    X = A + to.EPS
    v = X.var(axis=0, ddof=1)
    yc = y - y.mean()
    for i, j in rng.integers(0, l, size=(300, 2)):
        if v[i] < 0.01 or v[j] < 0.01:
            assert B[i, j] == 0.0
            continue
        z = X[:, i] * X[:, j]
        zc = z - z.mean()
        assert abs(B[i, j] - (zc @ yc) / (zc @ zc)) < RTOL * max(1.0, abs(B[i, j]))
    # the planted interaction (loci 11 and 701, effect 4 on the product) is recovered
    z = X[:, 10] * X[:, 700]
    zc = z - z.mean()
    assert abs(B[10, 700] - (zc @ yc) / (zc @ zc)) < RTOL * abs(B[10, 700]) and 3.0 < B[10, 700] < 5.5
    assert np.all(np.diff(cnt) > 0) and np.all(np.abs(full[cnt - 1]) >= np.sort(np.abs(full))[-1000])
    dm.free()


def test_host_mirror_matches_the_oracle_end_to_end(gbm):
    """transform1 / transform2 / epistasisfeatures with Genomes / Phenomes in, Genomes out: names,
    order and values as the reference builds them."""
    from gbm_b200 import transform as tr

    n, l = 80, 24
    A = synth.block(4, n, 0, l, synth.KIND_TETRAPLOID)
    y = synth.phenotype(4, n, l, synth.KIND_TETRAPLOID, n_causal=3)
    names = [f"chr1\t{j + 1}\tA|T\tA" for j in range(l)]
    G = gbm.Genomes.from_matrix(A, loci_alleles=names)
    P = gbm.Phenomes.from_matrix(y)
    out = tr.transform1(tr.square, G, P, n_new_features_per_transformation=7)
    _, idx, T = to.transform1(to.square, A, y, n_new=7)
    assert out.loci_alleles == [f"square({names[j - 1]})" for j in idx]
    assert np.array_equal(out.allele_frequencies, T)
    out2 = tr.transform2(tr.raise_, G, P, n_new_features_per_transformation=9)
    _, _, pairs, T2 = to.transform2(to.raise_, A, y, n_new=9)
    assert out2.loci_alleles == [f"raise({names[i - 1]},{names[j - 1]})" for i, j in pairs]
    np.testing.assert_allclose(out2.allele_frequencies, T2, rtol=4e-16, atol=0)
    ef = tr.epistasisfeatures(G, P, n_new_features_per_transformation=5, n_reps=2)
    B, new_names = to.epistasisfeatures(A, y, names, n_new=5, n_reps=2)
    assert ef.loci_alleles == new_names
    np.testing.assert_allclose(ef.allele_frequencies, B, rtol=1e-12, atol=1e-15)
    assert ef.checkdims()


def test_error_behaviour(gbm):
    from gbm_b200 import transform as tr

    A = synth.block(4, 30, 0, 6, synth.KIND_CONTINUOUS)
    y = np.arange(30.0)
    G, P = gbm.Genomes.from_matrix(A), gbm.Phenomes.from_matrix(y)
    with pytest.raises(gbm.ArgumentError):  # arbitrary closures cannot run on the device; no CPU fallback
        tr.transform1(lambda x: x ** 2, G, P, n_new_features_per_transformation=3)
    with pytest.raises(gbm.ArgumentError):
        tr.transform2(tr.square, G, P, n_new_features_per_transformation=3)
    with pytest.raises(gbm.ArgumentError, match="BoundsError"):  # sortperm(...)[1:n_new] on 6 effects
        tr.transform1(tr.square, G, P)
    Aneg = A.copy()
    Aneg[:, 2] = -Aneg[:, 2] - 0.5
    with pytest.raises(gbm.ArgumentError, match="Cannot transform"):  # log10 of a negative number: DomainError in Julia
        tr.transform1(tr.log10epsdivlog10eps, gbm.Genomes.from_matrix(Aneg), P, n_new_features_per_transformation=3)
    out = tr.transform1(tr.log10epsdivlog10eps, gbm.Genomes.from_matrix(Aneg), P, n_new_features_per_transformation=3,
                        use_abs=True)
    assert out.checkdims()


def test_packed_matrix_gives_the_same_effects(gbm):
    tr, _, _ = _pairs(gbm)
    A = synth.block(8, 300, 0, 100, synth.KIND_DIPLOID)
    y = synth.phenotype(8, 300, 100, synth.KIND_DIPLOID, n_causal=5)
    dm = gbm.DeviceMatrix.upload(A)
    pk = dm.pack()
    a, ia = tr.transform1_screen(dm, y, tr.invoneplus, 50)
    b, ib = tr.transform1_screen(pk, y, tr.invoneplus, 50)
    assert np.array_equal(a, b) and np.array_equal(ia, ib)
    a2, ca, _ = tr.transform2_screen(dm, y, tr.addnorm, 50, want_beta=True)
    b2, cb, _ = tr.transform2_screen(pk, y, tr.addnorm, 50, want_beta=True)
    assert np.array_equal(a2, b2) and np.array_equal(ca, cb)
    dm.free()
    pk.free()


@pytest.mark.parametrize("commutative", [False, True])
def test_row_sharded_screen_equals_the_unsharded_one(gbm, commutative):
    """gbm_transform2_screen_rows is one rank's share of a multi-GPU screen (a block of rows of the l x l pair
    matrix, no data-path collective): merging the per-block candidate lists must reproduce the unsharded
    selection bit for bit, for tile-aligned and ragged row blocks."""
    tr, _, _ = _pairs(gbm)
    n, l, n_new = 500, 333, 200
    A = synth.block(17, n, 0, l, synth.KIND_TETRAPLOID)
    y = synth.phenotype(17, n, l, synth.KIND_TETRAPLOID, n_causal=5)
    dm = gbm.DeviceMatrix.upload(A)
    beta, cnt, vals = tr.transform2_screen(dm, y, tr.addnorm, n_new, commutative=commutative, want_beta=True)
    for bounds in ([0, 128, 256, l], [0, 100, 101, 250, l], [0, l]):
        parts = [tr.transform2_screen_rows(dm, y, tr.addnorm, r0, r1, n_new, commutative=commutative)
                 for r0, r1 in zip(bounds[:-1], bounds[1:])]
        for (c, v), r0, r1 in zip(parts, bounds[:-1], bounds[1:]):
            assert np.all((c - 1) // l >= r0) and np.all((c - 1) // l < r1)
            assert np.array_equal(v, beta[c - 1])
            assert np.all(np.abs(v[:-1]) >= np.abs(v[1:]))  # selection order
        mc, mv = tr.merge_screen_candidates(parts, n_new)
        assert np.array_equal(mc, cnt) and np.array_equal(mv, vals)
    assert np.array_equal(tr.transform2_screen_sharded(dm, y, tr.addnorm, n_new, commutative=commutative)[0], cnt)
    dm.free()
