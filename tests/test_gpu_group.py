"""Multi-GPU entry points of the C ABI (gbm_group / gbm_sharded, include/gbm_b200.h): parity of the sharded
pipeline with the CPU oracle and with the single-GPU entry points.  Every test runs with a group of ONE GPU
(NCCL with a single rank: the whole code path, no exchange partner) so that the driver's one-GPU `-m gpu` run
covers it; the `multigpu` cases add 2 (and all visible) GPUs when the box has them (`gpurun --gpus N`)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import gwas_oracle as go, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def visible_gpus():
    import ctypes

    try:
        rt = ctypes.CDLL("libcudart.so.12")
        c = ctypes.c_int()
        return c.value if rt.cudaGetDeviceCount(ctypes.byref(c)) == 0 else 0
    except OSError:
        return 0


def group_sizes():
    n = visible_gpus()
    sizes = [1] + [g for g in (2, 4, 8) if g <= n]
    return [pytest.param(g, marks=[pytest.mark.multigpu] if g > 1 else []) for g in sizes]


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.fixture(scope="module", params=group_sizes())
def grp(request, gbm):
    from gbm_b200 import multigpu

    g = multigpu.Group.local(request.param)
    assert (g.world, g.n_local, g.first_rank) == (request.param, request.param, 0)
    yield g
    g.free()


@pytest.mark.parametrize("kind,n,p", [(synth.KIND_DIPLOID, 300, 1003), (synth.KIND_CONTINUOUS, 257, 640)])
def test_sharded_colstats_and_scan_equal_the_single_gpu_entry_points(gbm, grp, kind, n, p):
    from gbm_b200 import multigpu

    A = synth.block(11, n, 0, p, kind)
    y = synth.phenotype(11, n, p, kind)
    ys = (y - y.mean()) / y.std(ddof=1)
    c = np.random.default_rng(3).normal(size=(n, 1))
    sm = multigpu.ShardedMatrix.upload(grp, A, compact=(kind != synth.KIND_CONTINUOUS))
    assert (sm.n, sm.p) == (n, p) and sm.packed == (kind != synth.KIND_CONTINUOUS)
    assert sum(sm.ncols) == p and sm.first_col[0] == 0
    dm = gbm.DeviceMatrix.upload(A)
    if sm.packed:  # same storage on both sides: the comparison below is then bit for bit
        f64, dm = dm, dm.pack()
        f64.free()
    try:
        s1, s0 = sm.colstats(), dm.colstats()
        for k in ("mean", "sd", "min_nonzero", "keep", "idx_cols"):
            assert np.array_equal(s1[k], s0[k], equal_nan=True), k  # per-marker results do not depend on the sharding
        assert s1["min_nonzero_kept"] == s0["min_nonzero_kept"]
        r1 = sm.scan(np.c_[ys, ys[::-1]], c, model=gbm._lib.MODEL_LMM)
        r0 = dm.scan(np.c_[ys, ys[::-1]], c, model=gbm._lib.MODEL_LMM)
        for k in ("beta", "se", "stat", "neglog10p", "mean", "sd", "keep"):
            assert np.array_equal(r1[k], r0[k], equal_nan=True), k
    finally:
        sm.free()
        dm.free()


@pytest.mark.parametrize("grm_type,kind,n,p", [("simple", synth.KIND_DIPLOID, 300, 2049),
                                               ("ploidy-aware", synth.KIND_TETRAPLOID, 512, 3000),
                                               ("simple", synth.KIND_CONTINUOUS, 1100, 1500)])
def test_sharded_grm_and_pc1_match_the_oracle(gbm, grp, grm_type, kind, n, p):
    from gbm_b200 import multigpu

    A = synth.block(5, n, 0, p, kind)
    sm = multigpu.ShardedMatrix.upload(grp, A)
    try:
        if grm_type == "simple":
            K, tf = sm.grm(gbm._lib.GRM_SIMPLE)
            Kref = go.grm_simple(A)
        else:
            K, tf = sm.grm(gbm._lib.GRM_PLOIDY_AWARE, ploidy=4)
            Kref = go.grm_ploidy_aware(A, 4)
        assert np.max(np.abs(K - Kref)) <= 1e-9 * np.abs(Kref).max()
        assert np.array_equal(K, K.T) and tf > 0
        pc, _ = sm.kstd_pc1()                       # from the GRM left resident on the GPUs
        pc_ref = go.pca_pc1(go.standardise_K(Kref))
        if pc @ pc_ref < 0:
            pc = -pc
        assert abs(np.linalg.norm(pc) - 1) < 1e-12 and abs(pc.sum()) < 1e-9
        assert np.max(np.abs(pc - pc_ref)) < 1e-9
        pc2, _ = sm.kstd_pc1(Kref)                  # from a host GRM
        if pc2 @ pc_ref < 0:
            pc2 = -pc2
        assert np.max(np.abs(pc2 - pc_ref)) < 1e-9
    finally:
        sm.free()


def test_sharded_lanczos_pc1_over_peer_memory_and_over_nccl(gbm, grp, monkeypatch):
    """n >= 4,096: the columns of K are sharded and every Lanczos step ends in an all-reduce of an n-vector -- one kernel
    over peer memory (peer_sum_kernel: NVLink stores into every GPU's mailbox, rank-order sum) or, with
    GBM_PC1_PEER=0, NCCL.  Both must give the single-GPU routine's vector (gwas.jl:234)."""
    from gbm_b200 import multigpu

    n, p = 4224, 4096
    A = synth.block(17, n, 0, p, synth.KIND_DIPLOID)
    K = go.grm_simple(A)
    _, pc_one, _ = gbm.kstd_pc1(K, want_kstd=False)
    sm = multigpu.ShardedMatrix.upload(grp, A[:, :64])
    try:
        got = {}
        for route in ("1", "0", "1"):
            monkeypatch.setenv("GBM_PC1_PEER", route)
            pc, ms = sm.kstd_pc1(K)
            assert ms > 0 and abs(np.linalg.norm(pc) - 1) < 1e-12
            assert min(np.max(np.abs(pc - pc_one)), np.max(np.abs(pc + pc_one))) < 1e-9
            if route in got:  # the same route twice: the same bits (fixed-order sums, no atomics)
                assert np.array_equal(pc, got[route])
            got[route] = pc
        assert min(np.max(np.abs(got["1"] - got["0"])), np.max(np.abs(got["1"] + got["0"]))) < 1e-10
    finally:
        sm.free()


@pytest.mark.parametrize("model,grm_type,kind,n,p", [("lmm", "simple", synth.KIND_DIPLOID, 400, 3001),
                                                     ("ols", "ploidy-aware", synth.KIND_TETRAPLOID, 300, 2000),
                                                     ("lmm", "simple", synth.KIND_CONTINUOUS, 1200, 900)])
def test_one_call_gwas_matches_the_oracle(gbm, grp, model, grm_type, kind, n, p):
    """gbm_sharded_gwas == oracle gwasols / gwaslmm (restatement of /root/reference/src/gwas.jl:206-259, :329-399)
    at 1e-9 relative, filter indices exact."""
    from gbm_b200 import multigpu

    A = synth.block(23, n, 0, p, kind)
    y = synth.phenotype(23, n, p, kind)
    ent = [str(i) for i in range(n)]
    fn = go.gwaslmm if model == "lmm" else go.gwasols
    want, prep, _ = fn(A, ent, y[:, None], ent, GRM_type=grm_type)
    ys = (y - y.mean()) / y.std(ddof=1)
    sm = multigpu.ShardedMatrix.generate(grp, 23, n, p, kind, pack=True)
    try:
        res = sm.gwas(ys, model=gbm._lib.MODEL_LMM if model == "lmm" else gbm._lib.MODEL_OLS,
                      grm_type=gbm._lib.GRM_PLOIDY_AWARE if grm_type == "ploidy-aware" else gbm._lib.GRM_SIMPLE)
    finally:
        sm.free()
    assert np.array_equal(res["idx_cols"], prep.idx_cols)
    got = res["stat"][res["idx_cols"] - 1]
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-4 * np.abs(want).max())) < 1e-9
    tm = res["timing"]
    if grm_type == "ploidy-aware":
        assert tm["ploidy"] == 4
    assert tm["total_ms"] > 0 and tm["launches"] > 0 and tm["grm_tflops"] > 0


def test_group_of_ranks_with_one_rank(gbm):
    """gbm_group_create_rank (one process per GPU) with a world of one: the id hand-off and the inline path."""
    from gbm_b200 import multigpu

    uid = multigpu.Group.unique_id()
    assert len(uid) == 128
    g = multigpu.Group.from_rank(uid, 1, 0)
    try:
        n, p = 300, 777
        blocks = [gbm.DeviceMatrix.generate(9, n, p, synth.KIND_DIPLOID)]
        sm = multigpu.ShardedMatrix.adopt(g, blocks)
        assert (sm.n, sm.p, sm.ncols, sm.first_col) == (n, p, [p], [0])
        A = synth.block(9, n, 0, p, synth.KIND_DIPLOID)
        y = synth.phenotype(9, n, p, synth.KIND_DIPLOID)
        ent = [str(i) for i in range(n)]
        want, prep, _ = go.gwaslmm(A, ent, y[:, None], ent, GRM_type="simple")
        res = sm.gwas((y - y.mean()) / y.std(ddof=1))
        sm.free()
        blocks[0].free()
        assert np.array_equal(res["idx_cols"], prep.idx_cols)
        got = res["stat"][res["idx_cols"] - 1]
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-4 * np.abs(want).max())) < 1e-9
    finally:
        g.free()


def test_group_argument_errors(gbm):
    from gbm_b200 import multigpu

    with pytest.raises(gbm._lib.ArgumentError):
        multigpu.Group.local(0)
    with pytest.raises(gbm._lib.ArgumentError):
        multigpu.Group.local(visible_gpus() + 1)
    g = multigpu.Group.local(1)
    try:
        with pytest.raises(gbm._lib.ArgumentError):
            multigpu.ShardedMatrix.upload(g, np.zeros((1, 5)))
        sm = multigpu.ShardedMatrix.generate(g, 1, 64, 100, synth.KIND_DIPLOID)
        with pytest.raises(gbm._lib.ArgumentError):  # GRM_type check of gwasprep (gwas.jl:101-107)
            sm.grm(7)
        with pytest.raises(gbm._lib.ArgumentError):  # no resident GRM yet
            sm.kstd_pc1()
        sm.free()
    finally:
        g.free()


@pytest.mark.parametrize("n_gpus", group_sizes())
def test_c_program_drives_the_group_through_the_abi_alone(gbm, n_gpus):
    """A C99 program (tests/c/group_driver.c) links against libgbm_b200.so and runs the whole sharded gwaslmm on a
    local group -- no Python, no torch.distributed in that process."""
    import gbm_b200

    lib = gbm_b200.build()
    n, p, seed, kind = 384, 2500, 31, synth.KIND_TETRAPLOID
    y = synth.phenotype(seed, n, p, kind)
    ys = (y - y.mean()) / y.std(ddof=1)
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "group_driver")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "c", "group_driver.c"), "-o", exe, lib,
                               "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath,/usr/local/cuda/lib64"])
        out = subprocess.run([exe, str(n_gpus), str(n), str(p), str(seed), str(kind), "1", "1"], input=ys.tobytes(),
                             capture_output=True)
    assert out.returncode == 0, (out.returncode, out.stderr.decode()[-2000:])
    lines = [s.split()[1:] for s in out.stdout.decode().split("\n") if s.startswith("gbm ")]
    l, ploidy, packed = (int(v) for v in lines[0])
    assert len(lines) == 1 + l
    idx = np.array([int(s[0]) for s in lines[1:]], dtype=np.int64)
    z = np.array([float(s[1]) for s in lines[1:]])
    A = synth.block(seed, n, 0, p, kind)
    ent = [str(i) for i in range(n)]
    want, prep, _ = go.gwaslmm(A, ent, y[:, None], ent, GRM_type="ploidy-aware")
    assert ploidy == 4 and packed == 1
    assert np.array_equal(idx, prep.idx_cols)
    assert np.max(np.abs(z - want) / np.maximum(np.abs(want), 1e-4 * np.abs(want).max())) < 1e-9


@pytest.mark.multigpu
@pytest.mark.skipif(visible_gpus() < 2, reason="needs at least 2 GPUs on the box")
def test_one_process_per_gpu_under_torchrun(gbm):
    """tests/dist_gpu_check.py on every visible GPU: the RANK group (gbm_group_create_rank) against the oracle."""
    import sys

    import signal

    n = visible_gpus()
    port = 29600 + os.getpid() % 300
    # own session: if a rank dies the others wait in NCCL for ever -- the whole process group is killed at the timeout
    proc = subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                             "--master-addr", "127.0.0.1", "--master-port", str(port),
                             os.path.join(ROOT, "tests", "dist_gpu_check.py")], stdout=subprocess.PIPE,
                            stderr=subprocess.STDOUT, text=True, start_new_session=True)
    try:
        log, _ = proc.communicate(timeout=300)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        log, _ = proc.communicate()
        log += "\n[killed at the 300 s timeout]"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dist_gpu_check_{n}gpu.log"), "w") as f:
        f.write(log)
    assert proc.returncode == 0, log[-3000:]
    assert log.count("-> OK") == 2 and "FAIL" not in log, log[-3000:]


@pytest.mark.parametrize("n_gpus", group_sizes())
def test_host_mirror_uses_the_group_when_gbm_num_gpus_is_set(gbm, n_gpus, monkeypatch):
    """gwasols / gwaslmm of the host mirror (same keyword API as /root/reference/src/gwas.jl:206-214, :329-337) with
    GBM_NUM_GPUS: the Fit equals the single-GPU one."""
    from gbm_b200 import multigpu

    n, p = 320, 1500
    A = synth.block(3, n, 0, p, synth.KIND_TETRAPLOID)
    y = synth.phenotype(3, n, p, synth.KIND_TETRAPLOID)
    g = gbm.Genomes.from_matrix(A)
    ph = gbm.Phenomes.from_matrix(y, entries=g.entries)
    monkeypatch.delenv("GBM_NUM_GPUS", raising=False)
    one = gbm.gwaslmm(genomes=g, phenomes=ph, GRM_type="ploidy-aware")
    monkeypatch.setenv("GBM_NUM_GPUS", str(max(n_gpus, 2) if visible_gpus() >= 2 else 1))
    if visible_gpus() < 2:  # a single visible GPU: exercise the group path with a group of one
        monkeypatch.setattr(multigpu, "default_group", lambda: multigpu.Group.local(1))
    many = gbm.gwaslmm(genomes=g, phenomes=ph, GRM_type="ploidy-aware")
    assert many.model == one.model == "GWAS_LMM" and many.b_hat_labels == one.b_hat_labels
    assert "n_gpus" in many.extras and many.extras["ploidy"] == 4
    np.testing.assert_allclose(many.b_hat, one.b_hat, rtol=1e-9, atol=1e-9 * np.abs(one.b_hat).max())
