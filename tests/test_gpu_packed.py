"""GPU parity tests for the compact (one byte per genotype) storage path (SURVEY.md 8f
rank 3): every result must equal the Float64 path's on the same data."""
import numpy as np
import pytest

from oracle import gwas_oracle as go, synth

pytestmark = pytest.mark.gpu


def _problem(seed, n, p, kind):
    A = synth.block(seed, n, 0, p, kind)
    y = synth.phenotype(seed, n, p, kind)
    ys = (y - y.mean()) / y.std(ddof=1)
    rng = np.random.default_rng(seed)
    pc = rng.normal(size=n)
    return A, ys, pc


@pytest.mark.parametrize("n,p,kind", [(300, 1000, synth.KIND_TETRAPLOID), (1025, 333, synth.KIND_DIPLOID),
                                      (7, 5, synth.KIND_DIPLOID), (2049, 64, synth.KIND_TETRAPLOID),
                                      (10000, 257, synth.KIND_DIPLOID)])
def test_packed_scan_equals_float64_scan(gbm, n, p, kind):
    A, ys, pc = _problem(3, n, p, kind)
    dm = gbm.DeviceMatrix.upload(A)
    pk = dm.pack()
    assert pk is not None and pk.packed
    assert np.array_equal(pk.download(), A)  # decode is exact
    st_f, st_p = dm.colstats(), pk.colstats()
    assert np.array_equal(st_f["idx_cols"], st_p["idx_cols"])
    assert np.array_equal(st_f["keep"], st_p["keep"])
    assert st_f["min_nonzero_kept"] == st_p["min_nonzero_kept"]
    np.testing.assert_allclose(st_p["mean"], st_f["mean"], rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(st_p["sd"], st_f["sd"], rtol=1e-13, atol=1e-16)
    for model in (0, 1):
        a = dm.scan(ys, pc[:, None], model=model)
        b = pk.scan(ys, pc[:, None], model=model)
        keep = a["keep"]
        assert np.array_equal(keep, b["keep"])
        for key in ("beta", "se", "stat"):
            scale = np.nanmax(np.abs(a[key][keep]))
            assert np.nanmax(np.abs(a[key][keep] - b[key][keep])) < 1e-11 * max(scale, 1e-300), key
        assert np.nanmax(np.abs(a["neglog10p"][keep] - b["neglog10p"][keep])) < 1e-9
        assert np.all(np.isnan(b["stat"][~keep]))
    # oracle, directly
    mu, v = go.column_std(A)
    keep = v > go.EPS
    if n > 4:
        pcn = pc - pc.mean()
        ref = go.scan_closed_form(A[:, keep], ys, pcn / np.linalg.norm(pcn))
        got = pk.scan(ys, pc[:, None], model=1)["stat"][keep, 0]
        assert np.max(np.abs(got - ref["stat_lmm"]) / np.maximum(np.abs(ref["stat_lmm"]), 1e-4 * np.abs(ref["stat_lmm"]).max())) < 1e-9
    dm.free()
    pk.free()


def test_pack_refuses_inexact_values(gbm):
    A = synth.block(3, 200, 0, 50, synth.KIND_CONTINUOUS)  # k/4096 grid: not all multiples of 1/240
    dm = gbm.DeviceMatrix.upload(A)
    assert dm.pack() is None
    dm.free()
    B = synth.block(3, 200, 0, 50, synth.KIND_DIPLOID)
    B[17, 3] = 0.3000000001
    dm = gbm.DeviceMatrix.upload(B)
    assert dm.pack() is None
    dm.free()
    codes, bad = gbm.pack_host(B)
    assert bad == 1
    codes, bad = gbm.pack_host(synth.block(3, 200, 0, 50, synth.KIND_TETRAPLOID))
    assert bad == 0 and set(np.unique(codes)) <= {0, 60, 120, 180, 240}
    # thirds and sixths are codes too (hexaploid levels), as the doubles k/6
    H = np.asfortranarray(np.random.default_rng(0).integers(0, 7, size=(50, 20)) / 6.0)
    codes, bad = gbm.pack_host(H)
    assert bad == 0
    pk = gbm.DeviceMatrix.upload_packed(codes)
    assert np.array_equal(pk.download(), H)
    pk.free()


def test_packed_multi_trait_and_grm_and_lmm(gbm):
    from oracle import lmm_oracle as lo

    n, p = 300, 700
    A = synth.block(9, n, 0, p, synth.KIND_TETRAPLOID)
    rng = np.random.default_rng(1)
    Y = rng.normal(size=(n, 5))
    C = rng.normal(size=(n, 2))
    dm = gbm.DeviceMatrix.upload(A)
    pk = dm.pack()
    a = dm.scan(Y, C, model=1)
    b = pk.scan(Y, C, model=1)  # 7 side vectors: decoded blocks + Float64 kernel
    keep = a["keep"]
    assert np.nanmax(np.abs(a["stat"][keep] - b["stat"][keep])) < 1e-11
    K1, _ = dm.grm(1, 4, 0)
    K2, _ = pk.grm(1, 4, 0)
    assert np.max(np.abs(K1 - K2)) < 1e-13 * np.abs(K1).max()
    g = (A - A.mean(axis=0)) @ rng.normal(size=p)
    y = g / g.std() + rng.normal(size=n)
    plan = gbm.LmmPlan(go.grm_simple(A), y)
    r1 = plan.run(dm)
    r2 = plan.run(pk)
    plan.free()
    assert np.nanmax(np.abs(r1["stat"] - r2["stat"])) < 1e-10
    dm.free()
    pk.free()


@pytest.mark.parametrize("kind,expect_packed", [(synth.KIND_DIPLOID, True), (synth.KIND_CONTINUOUS, False)])
def test_scan_host_pack_flag(gbm, kind, expect_packed):
    n, p = 1000, 9000
    A, ys, pc = _problem(13, n, p, kind)
    a = gbm.scan_host(A, ys, pc[:, None], model=1, pack=False)
    b = gbm.scan_host(A, ys, pc[:, None], model=1)  # auto-pack is the default
    tm = gbm.last_timing()
    # a block that is all codes is scanned as codes whichever lane carried it
    assert (tm["packed_blocks"] > 0) == expect_packed
    assert tm["host_packed_blocks"] <= tm["packed_blocks"]
    keep = a["keep"]
    assert np.array_equal(keep, b["keep"])
    tol = 1e-11 if expect_packed else 0.0
    for key in ("beta", "se", "stat", "mean", "sd"):
        x, y = a[key][keep], b[key][keep]
        assert np.nanmax(np.abs(x - y)) <= tol * max(1.0, np.nanmax(np.abs(x))), key


def test_scan_host_lanes_agree_bit_for_bit_on_mixed_data(gbm, monkeypatch):
    """gbm_scan_host hands ~128 MB column blocks to the host-packer lane and/or the copy-engine lane
    (GBM_SCAN_HOST_LANES).  A block that is all dosage codes is scanned as codes, any other block as
    Float64, whichever lane carried it -- so every lane configuration must give the same bits, also when
    the host packer meets a non-code block half way (hand-back to the copy-engine lane) and when the
    source is page-locked."""
    import torch

    n, p = 1000, 56_000  # 3.5 blocks of 16,640 columns (n off the 16-row pitch: blocks are re-pitched)
    A, ys, pc = _problem(21, n, p, synth.KIND_DIPLOID)
    C = synth.block(22, n, 0, 3000, synth.KIND_CONTINUOUS)
    A[:, 35_000:38_000] = C  # the third block is not dosage data
    pinned = torch.empty((p, n), dtype=torch.float64, pin_memory=True)
    pinned.numpy()[:] = A.T
    ref = None
    seen = set()
    for src in (A, pinned):
        for lanes in (None, "host", "copy", "both"):
            if lanes is None:
                monkeypatch.delenv("GBM_SCAN_HOST_LANES", raising=False)
            else:
                monkeypatch.setenv("GBM_SCAN_HOST_LANES", lanes)
            out = gbm.scan_host(src, ys, pc[:, None], model=1)
            tm = gbm.last_timing()
            assert tm["packed_blocks"] == 3  # blocks 1, 2 and the ragged 4th are codes, block 3 is not
            seen.add(tm["host_packed_blocks"])
            if ref is None:
                ref = out
                continue
            for key in ("beta", "se", "stat", "neglog10p", "mean", "sd", "keep"):
                assert np.array_equal(ref[key], out[key], equal_nan=True), (lanes, key)
    monkeypatch.delenv("GBM_SCAN_HOST_LANES", raising=False)
    assert 0 in seen and max(seen) >= 2  # both the copy-engine-only and the host-lane paths ran
    dm = gbm.DeviceMatrix.upload(A)
    want = dm.scan(ys, pc[:, None], model=1)
    dm.free()
    keep = want["keep"]
    assert np.array_equal(keep, ref["keep"])
    for key in ("beta", "se", "stat"):
        x, y = want[key][keep], ref[key][keep]
        assert np.nanmax(np.abs(x - y)) <= 1e-11 * max(1.0, np.nanmax(np.abs(x))), key


@pytest.mark.parametrize("n,p", [(1000, 9000), (1008, 5000), (37, 11), (4097, 2100)])
def test_upload_paths_roundtrip(gbm, n, p):
    """gbm_matrix_upload from pageable memory (host workers -> pinned staging ring -> copy engine) and from
    page-locked memory (copy engine directly), with and without re-pitching: bit-exact round trips."""
    import torch

    A = synth.block(31, n, 0, p, synth.KIND_CONTINUOUS)
    dm = gbm.DeviceMatrix.upload(A)
    assert np.array_equal(dm.download(), A)
    dm.free()
    pinned = torch.empty((p, n), dtype=torch.float64, pin_memory=True)
    pinned.numpy()[:] = A.T
    m = gbm.DeviceMatrix.upload_compact(pinned)  # continuous data: Float64 slab
    assert not m.packed and np.array_equal(m.download(), A)
    m.free()


@pytest.mark.parametrize("kind,packs", [(synth.KIND_TETRAPLOID, True), (synth.KIND_DIPLOID, True),
                                        (synth.KIND_CONTINUOUS, False)])
def test_upload_compact(gbm, kind, packs):
    """The host cores pack on the way to the device when every element is a dosage code; the handle then
    behaves like the device-packed one (same codes, same scan results, Float64 downloads)."""
    n, p = 1500, 20_000  # two 128 MB-equivalent blocks
    A, ys, pc = _problem(5, n, p, kind)
    m = gbm.DeviceMatrix.upload_compact(A)
    assert m.packed == packs
    assert np.array_equal(m.download(), A)
    dm = gbm.DeviceMatrix.upload(A)
    ref = dm.pack() if packs else dm
    a = ref.scan(ys, pc[:, None], model=1)
    b = m.scan(ys, pc[:, None], model=1)
    for key in ("beta", "se", "stat", "neglog10p", "mean", "sd", "keep"):
        assert np.array_equal(a[key], b[key], equal_nan=True), key
    # G[:, idx_cols] standardised on the device, also from codes (gwas.jl:114, :129)
    idx = np.flatnonzero(a["keep"])[:700] + 1
    G1 = dm.download_cols(idx, standardise=True)
    G2 = m.download_cols(idx, standardise=True)
    np.testing.assert_allclose(G2, G1, rtol=1e-12, atol=1e-13)
    want = (A[:, idx - 1] - A[:, idx - 1].mean(axis=0)) / A[:, idx - 1].std(axis=0, ddof=1)
    np.testing.assert_allclose(G2, want, rtol=1e-10, atol=1e-12)
    if packs:
        ref.free()
        # a single element that is not a code anywhere in the matrix: Float64 slab, nothing lost
        B = A.copy()
        B[n - 1, p - 1] = 0.123
        mb = gbm.DeviceMatrix.upload_compact(B)
        assert not mb.packed and np.array_equal(mb.download(), B)
        mb.free()
    dm.free()
    m.free()


@pytest.mark.parametrize("kind", [synth.KIND_TETRAPLOID, synth.KIND_CONTINUOUS])
def test_config5_reduced_multi_trait_from_host(gbm, kind):
    """BASELINE configs[4] reduced (20 traits, host-resident matrix streamed in blocks): gbm_scan_host with
    T = 20 and one covariate needs two passes of side vectors per block (13 + 7 traits); every trait column
    must equal the resident multi-trait scan and the oracle's closed form."""
    n, p, T = 1200, 30_000, 20
    A = synth.block(8, n, 0, p, kind)
    rng = np.random.default_rng(8)
    Y = rng.normal(size=(n, T)) + A[:, :T] * 0.5
    pc = rng.normal(size=n)
    host = gbm.scan_host(A, Y, pc[:, None], model=0)
    dm = gbm.DeviceMatrix.upload(A)
    res = dm.scan(Y, pc[:, None], model=0)
    dm.free()
    keep = res["keep"]
    assert np.array_equal(keep, host["keep"])
    for key in ("beta", "se", "stat", "neglog10p"):
        x, y = res[key][keep], host[key][keep]
        assert x.shape[1] == T and np.nanmax(np.abs(x - y)) <= 1e-10 * max(1.0, np.nanmax(np.abs(x))), key
    sub = np.flatnonzero(keep)[:300]
    for t in (0, 12, 13, 19):  # both passes
        ys = Y[:, t]
        cf = go.scan_closed_form(A[:, sub], ys, pc)
        err = np.abs(host["stat"][sub, t] - cf["stat_ols"]) / np.maximum(np.abs(cf["stat_ols"]), 1e-4 * np.abs(cf["stat_ols"]).max())
        assert err.max() < 1e-9, (t, err.max())


@pytest.mark.parametrize("n,p", [(300, 500), (1000, 129), (10_000, 700), (131, 64), (70_001, 40)])
def test_tensor_core_code_scan_equals_the_cuda_core_kernel_and_the_oracle(gbm, n, p, monkeypatch):
    """scan_u8_tc.cu (dots as exact u8 x s8 integer contractions on the tcgen05 tensor cores, side vectors as seven
    base-256 digits) against the CUDA-core code kernel (GBM_U8_TC=0) and the oracle's closed form: ragged n and p
    (TMA zero fill), n beyond one s32 accumulation segment (65,536 rows), side vectors with a wide dynamic range."""
    rng = np.random.default_rng(n + p)
    A = synth.block(21, n, 0, p, synth.KIND_TETRAPLOID)
    A[:, 3] = 0.75  # constant code
    A[:, 4] = 0.0
    y = rng.normal(size=n) * np.exp(rng.normal(size=n) * 3.0)  # heavy tails: entries from 1e-4 to 1e4 of the scale
    pc = rng.normal(size=n)
    pc[rng.integers(0, n, 5)] *= 1e3
    dm = gbm.DeviceMatrix.upload(A)
    pk = dm.pack()
    assert pk is not None
    monkeypatch.setenv("GBM_U8_TC", "1")
    tc = pk.scan(y, pc[:, None], model=1)
    tc1 = pk.scan(y, None, model=0)  # one side vector only
    monkeypatch.setenv("GBM_U8_TC", "0")
    cc = pk.scan(y, pc[:, None], model=1)
    cc1 = pk.scan(y, None, model=0)
    monkeypatch.delenv("GBM_U8_TC")
    f64 = dm.scan(y, pc[:, None], model=1)
    dm.free()
    pk.free()
    keep = f64["keep"]
    assert np.array_equal(tc["keep"], keep) and np.array_equal(cc["keep"], keep) and not keep[3] and not keep[4]
    assert tc["sd"][3] == 0.0 and tc["sd"][4] == 0.0
    for key in ("mean", "sd"):  # integer sums in both code kernels: bit for bit
        assert np.array_equal(tc[key], cc[key]), key
    for a, b in ((tc, cc), (tc, f64), (tc1, cc1)):
        for key in ("beta", "se", "stat"):
            x, z = a[key][keep], b[key][keep]
            assert np.nanmax(np.abs(x - z)) <= 1e-10 * max(1.0, np.nanmax(np.abs(z))), key
    pcn = pc - pc.mean()
    ref = go.scan_closed_form(A[:, keep], y - y.mean(), pcn / np.linalg.norm(pcn))
    got = tc["stat"][keep, 0]
    assert np.max(np.abs(got - ref["stat_lmm"]) / np.maximum(np.abs(ref["stat_lmm"]), 1e-4 * np.abs(ref["stat_lmm"]).max())) < 1e-9
