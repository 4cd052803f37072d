"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/gbm_b200.h
declares; the host-side mirror validates arguments like the reference; compute calls fail
loudly (no fallback) when there is no GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gbm_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gbm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import gbm_b200

    path = gbm_b200.build()
    assert os.path.exists(path)
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (gbm_[a-z0-9_]+)", out))
    decl = declared_symbols()
    assert len(decl) >= 20
    assert set(decl) <= exported, sorted(set(decl) - exported)
    # the ctypes harness binds exactly the header's symbols
    assert sorted(gbm_b200._lib.SIGNATURES) == decl
    lib = gbm_b200.load()
    from gbm_b200 import _lib as L

    assert lib.gbm_abi_version() == L.ABI_VERSION == 3


def test_library_is_sm100a_with_tma_and_dmma():
    import gbm_b200

    sass = subprocess.run(["cuobjdump", "-sass", gbm_b200.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTMALDG" in sass and "DMMA" in sass and "SYNCS" in sass
    # the tcgen05 kernels (grm_i8.cu, scan_u8_tc.cu): kind::i8 MMA, tcgen05.ld, tcgen05.commit
    assert "UTCIMMA" in sass and "LDTM" in sass and "UTCBAR" in sass
    # sm_100a only: no other architecture in the fat binary
    import re

    elf = subprocess.run(["cuobjdump", "-lelf", gbm_b200.build()], capture_output=True, text=True).stdout
    assert set(re.findall(r"sm_\d+a?", elf)) == {"sm_100a"}


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    import gbm_b200

    with pytest.raises(gbm_b200.CudaError):
        gbm_b200.init(0)
    A = np.asfortranarray(np.random.default_rng(0).random((10, 5)))
    with pytest.raises(gbm_b200.CudaError):
        gbm_b200.DeviceMatrix.upload(A)
    g = gbm_b200.Genomes.from_matrix(A)
    ph = gbm_b200.Phenomes.from_matrix(np.arange(10.0), entries=g.entries)
    with pytest.raises(gbm_b200.CudaError):
        gbm_b200.gwasols(genomes=g, phenomes=ph)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genomicbreedingmodels.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f


def test_host_validation_mirrors_reference_errors():
    """Errors of extractxyetc / gwasprep raised before any device work
    (/root/reference/src/prediction.jl:67-127, /root/reference/src/gwas.jl:101-111)."""
    import gbm_b200
    from gbm_b200.gwas import _validate_and_select

    rng = np.random.default_rng(1)
    A = np.asfortranarray(rng.random((12, 7)))
    y = rng.normal(size=12)
    g = gbm_b200.Genomes.from_matrix(A)
    ph = gbm_b200.Phenomes.from_matrix(y, entries=g.entries)
    rows, cols, yy = _validate_and_select(g, ph, None, None, 1)
    assert rows is None and cols is None and np.array_equal(yy, y)
    # missing / NaN / Inf phenotypes are dropped (prediction.jl:116)
    y2 = y.copy()
    y2[[2, 5]] = np.nan
    y2[7] = -np.inf
    ph2 = gbm_b200.Phenomes.from_matrix(y2, entries=g.entries)
    rows, _, yy = _validate_and_select(g, ph2, None, None, 1)
    assert list(rows) == [1, 2, 4, 5, 7, 9, 10, 11, 12] and yy.size == 9
    rows, _, _ = _validate_and_select(g, ph2, [12, 3, 1, 6], None, 1)
    assert list(rows) == [12, 1]
    bad = gbm_b200.Phenomes.from_matrix(y, entries=g.entries[::-1])
    with pytest.raises(gbm_b200.ArgumentError, match="merged"):
        _validate_and_select(g, bad, None, None, 1)
    with pytest.raises(gbm_b200.ArgumentError, match="idx_entries"):
        _validate_and_select(g, ph, [0, 3], None, 1)
    with pytest.raises(gbm_b200.ArgumentError, match="idx_loci_alleles"):
        _validate_and_select(g, ph, None, [1, 8], 1)
    with pytest.raises(gbm_b200.ArgumentError, match="less than 2"):
        _validate_and_select(g, gbm_b200.Phenomes.from_matrix(np.full(12, np.nan), entries=g.entries), None, None, 1)
    with pytest.raises(gbm_b200.ErrorException, match="variance"):
        _validate_and_select(g, gbm_b200.Phenomes.from_matrix(np.ones(12), entries=g.entries), None, None, 1)
    corrupt = gbm_b200.Genomes.from_matrix(A)
    corrupt.entries = corrupt.entries[:-1]
    with pytest.raises(gbm_b200.ArgumentError, match="corrupted"):
        _validate_and_select(corrupt, ph, None, None, 1)
    # GRM_type is validated before the device is touched too (gwas.jl:101-107)
    with pytest.raises((gbm_b200.ArgumentError, gbm_b200.CudaError)):
        gbm_b200.gwasols(genomes=g, phenomes=ph, GRM_type="fancy")


def test_extractxyetc_mirror_identity():
    import gbm_b200

    rng = np.random.default_rng(2)
    A = np.asfortranarray(rng.random((9, 4)))
    y = rng.normal(size=9)
    g = gbm_b200.Genomes.from_matrix(A)
    ph = gbm_b200.Phenomes.from_matrix(y, entries=g.entries)
    X, yy, ent, pops, loci = gbm_b200.extractxyetc(g, ph)
    assert np.array_equal(X, np.hstack([np.ones((9, 1)), A])) and np.array_equal(yy, y)  # prediction.jl:46-50
    assert ent == g.entries and loci == g.loci_alleles and len(pops) == 9
    f = gbm_b200.Fit.new(9, 4)
    assert f.checkdims() and f.b_hat.shape == (4,) and np.all(f.b_hat == 0)


def test_pvalue_device_functions_on_host():
    """pvalue.cuh is __host__ __device__: compile it for the host and check the log-space
    tails against the oracle (mpmath) -- the same code the finalisation kernel runs."""
    import tempfile

    from oracle import gwas_oracle as go

    csrc = os.path.join(ROOT, "genomicbreedingmodels.jl_b200", "csrc")
    src = r"""
#include <cstdio>
#include <cstdlib>
#include "pvalue.cuh"
int main(int argc, char** argv) {
  double df = atof(argv[1]);
  for (int i = 2; i < argc; ++i) {
    double t = atof(argv[i]);
    printf("%.17g %.17g\n", -gbm::log_sf_t(t, df) / log(10.0), -gbm::log_sf_normal(t) / log(10.0));
  }
  return 0;
}
"""
    with tempfile.TemporaryDirectory() as td:
        cu = os.path.join(td, "pv.cu")
        open(cu, "w").write(src)
        exe = os.path.join(td, "pv")
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-ccbin", "/usr/bin/g++", "-O2", "-I", csrc, "-o", exe, cu])
        # includes both sides of the switch to Hill's expansion (df >= 200 and |t| <= 8)
        t = [0.0, 0.1, 0.5, 1.0, 1.7, 1.75, 2.0, 3.0, 5.0, 6.5, 7.99, 8.0, 8.01, 8.3, 12.0, 19.9, 20.1, 37.0, 40.0,
             100.0, 300.0]
        for df in (1.0, 4.0, 99.0, 199.0, 200.0, 299.0, 9999.0, 19999.0, 1e6):
            out = subprocess.run([exe, repr(df)] + [repr(x) for x in t], capture_output=True, text=True, check=True)
            got = np.array([[float(v) for v in line.split()] for line in out.stdout.strip().splitlines()])
            want_t = go.neglog10_sf_t(t, df)
            want_z = go.neglog10_sf_normal(t)
            assert np.max(np.abs(got[:, 0] - want_t)) < 1e-6, (df, np.abs(got[:, 0] - want_t).max())
            assert np.max(np.abs(got[:, 1] - want_z)) < 1e-6
            ok = want_t < 300
            assert np.max(np.abs(10.0 ** (want_t[ok] - got[ok, 0]) - 1.0)) < 1e-9, df


def test_host_packer_without_gpu():
    """gbm_pack_host is pure host code (AVX2 / scalar, thread pool): exactness check and codes."""
    import gbm_b200
    from oracle import synth

    for kind, levels in ((synth.KIND_DIPLOID, {0, 120, 240}), (synth.KIND_TETRAPLOID, {0, 60, 120, 180, 240})):
        A = synth.block(5, 1037, 0, 203, kind)  # n not a multiple of 8: scalar tail after the AVX2 body
        codes, bad = gbm_b200.pack_host(A)
        assert bad == 0 and set(np.unique(codes)) <= levels
        assert np.array_equal(codes.astype(np.float64) / 240.0, A)
    rng = np.random.default_rng(0)
    H = np.asfortranarray(rng.integers(0, 7, size=(333, 17)) / 6.0)  # hexaploid levels k/6 as doubles
    codes, bad = gbm_b200.pack_host(H)
    assert bad == 0 and np.array_equal(codes.astype(np.float64) / 240.0, H)
    for k in (3, 5, 8, 10, 12):
        L = np.asfortranarray(rng.integers(0, k + 1, size=(64, 9)) / float(k))
        assert gbm_b200.pack_host(L)[1] == 0
    # anything that is not exactly a code is counted
    B = synth.block(5, 500, 0, 40, synth.KIND_DIPLOID)
    B[3, 7] = 0.5000000000000001
    B[499, 39] = np.nan
    B[0, 0] = 1.5
    B[10, 10] = -0.5
    assert gbm_b200.pack_host(B)[1] == 4
    C = synth.block(5, 100, 0, 30, synth.KIND_CONTINUOUS)
    _, bad = gbm_b200.pack_host(C)
    want = int(np.sum((np.rint(C * 240.0) / 240.0) != C))
    assert bad == want and bad > 0


def test_host_packer_vector_predicate_equals_the_division():
    """The AVX2 / AVX-512 packer bodies decide `fl(c/240) == a` without dividing (host_pack.cpp);
    they must agree with the scalar division element for element: every code, its neighbouring
    doubles, half-way values, binade edges, specials and random doubles."""
    import ctypes

    from gbm_b200 import _lib

    lib = _lib.load()
    codes = np.arange(0, 241, dtype=np.float64) / 240.0
    cand = [codes]
    for k in (1, 2, 3, 7):
        up, dn = codes.copy(), codes.copy()
        for _ in range(k):
            up, dn = np.nextafter(up, 2.0), np.nextafter(dn, -1.0)
        cand += [up, dn]
    cand.append((np.arange(0, 240) + 0.5) / 240.0)
    cand.append(np.arange(0, 241, dtype=np.float64) * (1.0 / 240.0))  # c * fl(1/240): often 1 ulp off
    cand.append(2.0 ** -np.arange(0, 60.0))
    cand.append(np.array([-0.0, 5e-324, 1e-310, 2.2250738585072014e-308, 1e-300, -1e-300, np.nan, np.inf, -np.inf,
                          1.0000000000000002, 241.0 / 240.0, 2.0, 1e300, -1.0 / 240.0, 4.0e9, -4.0e9]))
    rng = np.random.default_rng(11)
    cand.append(rng.random(20000))
    cand.append(np.rint(rng.random(20000) * 240.0) / 240.0 + rng.integers(-2, 3, 20000) * 2.0 ** -54)
    v = np.concatenate(cand)
    A = np.asfortranarray(np.tile(v, (16, 1)))  # one column of 16 equal values per candidate: the vector body sees it
    n, p = A.shape
    got = {}
    for isa in (0, 1, 2):
        ok = np.zeros(p, dtype=np.uint8)
        rc = lib.gbm_pack_host_check(_lib.ptr(A), n, p, n, isa, _lib.ptr(ok))
        if rc != 0:
            assert isa > 0  # CPU without that ISA
            continue
        got[isa] = ok
    with np.errstate(all="ignore"):
        c = np.rint(v * 240.0)
        want = ((c >= 0) & (c <= 240) & (c / 240.0 == v)).astype(np.uint8)
    assert np.array_equal(got[0], want)
    assert len(got) >= 2, "no vector ISA on this host: nothing checked"
    for isa, ok in got.items():
        bad = np.nonzero(ok != want)[0]
        assert bad.size == 0, (isa, v[bad[:5]], ok[bad[:5]], want[bad[:5]])


def test_lanczos_tridiagonal_solver_on_the_host():
    """csrc/lanczos.cu solves the small tridiagonal problem on the host (bisection + inverse iteration with a
    pivoted solve): largest eigenpair against scipy on random, graded, nearly-degenerate and tiny cases."""
    import subprocess
    import tempfile

    from scipy.linalg import eigh_tridiagonal

    csrc = os.path.join(ROOT, "genomicbreedingmodels.jl_b200", "csrc")
    src = r"""
#include "lanczos.cu"
#include <stdio.h>
int main() {
  int m;
  while (scanf("%d", &m) == 1) {
    std::vector<double> a(m), b(m), s;
    for (int i = 0; i < m; ++i) scanf("%lf", &a[i]);
    for (int i = 0; i + 1 < m; ++i) scanf("%lf", &b[i]);
    double theta;
    gbm::tridiag_top(a, b, m, &theta, &s);
    printf("%.17g", theta);
    for (int i = 0; i < m; ++i) printf(" %.17g", s[i]);
    printf("\n");
  }
  return 0;
}
"""
    rng = np.random.default_rng(3)
    cases = []
    for m in (1, 2, 3, 7, 40, 333):
        cases.append((rng.normal(size=m) * 3 + 10, np.abs(rng.normal(size=max(m - 1, 0))) + 0.1))
    cases.append((np.linspace(1, 2, 50) ** 8, np.full(49, 1e-3)))           # graded
    cases.append((np.full(60, 5.0), np.full(59, 1e-9)))                      # nearly degenerate
    cases.append((np.array([4.0, 4.0, 1.0]), np.array([0.0, 0.5])))          # a zero coupling
    cases.append((rng.normal(size=200) * 1e3, rng.normal(size=199) * 1e3))   # negative couplings, large scale
    text = ""
    for a, b in cases:
        text += f"{a.size} " + " ".join(repr(float(x)) for x in a) + " " + " ".join(repr(float(x)) for x in b) + "\n"
    with tempfile.TemporaryDirectory() as td:
        cu = os.path.join(td, "tri.cu")
        open(cu, "w").write(src)
        exe = os.path.join(td, "tri")
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-ccbin", "/usr/bin/g++", "-O2", "-std=c++17",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-I", csrc, "-o", exe, cu])
        out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(cases)
    for (a, b), line in zip(cases, out):
        vals = np.array([float(v) for v in line.split()])
        theta, s = vals[0], vals[1:]
        if a.size == 1:
            assert theta == a[0] and s[0] == 1.0
            continue
        w, v = eigh_tridiagonal(a, b)
        scale = max(np.abs(w).max(), 1e-300)
        assert abs(theta - w[-1]) <= 4e-15 * scale, (a.size, theta, w[-1])
        T = np.diag(a) + np.diag(b, 1) + np.diag(b, -1)
        assert abs(np.linalg.norm(s) - 1.0) < 1e-12
        assert np.linalg.norm(T @ s - theta * s) <= 1e-13 * scale * np.sqrt(a.size), a.size


def test_graft_entry_build_runs_without_a_gpu():
    """The driver's build check: __graft_entry__.build() compiles (or finds up to date) every CUDA source
    for sm_100a, loads the library and builds the oracle's C twin."""
    import importlib

    entry = importlib.import_module("__graft_entry__")
    entry.build()


def test_public_header_is_plain_c_and_links_against_the_library():
    """include/gbm_b200.h is the drop-in boundary: it must compile as C99 (no C++-isms, no torch types) and a
    C program must link against libgbm_b200.so and call an entry point that needs no GPU."""
    import subprocess
    import tempfile

    import gbm_b200

    lib = gbm_b200.build()
    src = r"""
#include <stdio.h>
#include "gbm_b200.h"
int main(void) {
  double a[4] = {0.0, 0.5, 1.0, 0.25};
  unsigned char codes[4];
  long long bad = -1;
  int64_t inexact = -1;
  (void)bad;
  if (gbm_abi_version() != GBM_ABI_VERSION) return 2;
  if (gbm_pack_host(a, 4, 1, 4, codes, 4, &inexact) != GBM_OK) return 3;
  printf("%d %d %d %d %d\n", codes[0], codes[1], codes[2], codes[3], (int)inexact);
  return gbm_scan(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) == GBM_ERR_NOT_INITIALISED ? 0 : 4;
}
"""
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "abi.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "abi")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                               c, "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath,/usr/local/cuda/lib64"])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
        assert out.stdout.split() == ["0", "120", "240", "60", "0"]


def test_side_vector_digits_reconstruct_the_vectors_exactly():
    """The tensor-core code scan (csrc/scan_u8_tc.cu) multiplies the codes with FIXED-POINT side vectors: seven balanced
    base-256 digits per element.  Host-side check (no GPU): the digits are int8, the column of ones is there, and
    scale * sum_k 256^k d_k equals q to within one unit of the last digit -- below one ulp of max |q| -- for vectors with
    a wide dynamic range, values within a hair of a power of two (the carry that needs the headroom bit), zeros and
    both signs; an integer dot product with codes then equals the exact rational dot product to the same bound."""
    import ctypes
    from fractions import Fraction

    import gbm_b200
    from gbm_b200 import _lib as L

    lib = gbm_b200.load()
    rng = np.random.default_rng(0)
    n = 1000
    q1 = rng.normal(size=n) * np.exp(rng.normal(size=n) * 4.0)
    q1[0], q1[1], q1[2], q1[3] = 0.0, np.abs(q1).max() * 1.5, -np.abs(q1).max() * 1.4999, 1e-300
    q2 = rng.normal(size=n)
    q2[5] = np.nextafter(4.0, 0.0)   # just below a power of two: the top digit would carry without headroom
    q2[6] = -np.nextafter(4.0, 0.0)
    q2[7] = 3.999
    Q = np.asfortranarray(np.c_[q1, q2])
    ld = (n + 127) // 128 * 128
    dig = np.zeros((16, ld), dtype=np.int8)
    scale = (ctypes.c_double * 2)()
    L.check(lib.gbm_side_vector_digits(L.ptr(Q), n, 2, n, L.ptr(dig), ld, scale))
    assert np.all(dig[0, :n] == 1) and np.all(dig[0, n:] == 0) and np.all(dig[15] == 0) and np.all(dig[:, n:] == 0)
    for m, q in enumerate((q1, q2)):
        sc = scale[m]
        assert sc > 0 and np.log2(sc) == round(np.log2(sc))  # a power of two
        d = dig[1 + 7 * m: 8 + 7 * m, :n].astype(object)
        ints = sum(d[k] * (256 ** k) for k in range(7))      # exact Python integers
        assert max(abs(int(v)) for v in ints) < 2 ** 54
        err = np.array([abs(float(Fraction(int(v)) * Fraction(sc) - Fraction(float(x)))) for v, x in zip(ints, q)])
        assert err.max() <= 0.5 * sc * (1 + 1e-12), (m, err.max(), sc)       # round to nearest unit of the last digit
        assert sc <= 2.0 * np.spacing(np.abs(q).max())                        # that unit is within 2 ulp of max |q|
        codes = rng.integers(0, 241, size=n)
        exact = sum(Fraction(int(c)) * Fraction(float(x)) for c, x in zip(codes, q))
        fixed = Fraction(sc) * sum(int(c) * int(v) for c, v in zip(codes, ints))
        assert abs(float(fixed - exact)) <= 0.5 * sc * float(codes.sum())
    with pytest.raises(L.ArgumentError):
        L.check(lib.gbm_side_vector_digits(L.ptr(Q), n, 3, n, L.ptr(dig), ld, scale))


# ---- the Julia shim against the header (static: there is no Julia runtime in the image) ---------------------------
def _split_top_level(text):
    """Split on commas that are not nested in (), [] or {}."""
    parts, depth, cur = [], 0, []
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        parts.append(tail)
    return parts


def _balanced(text, start):
    """text[start] == '(' -> index just past its matching ')'."""
    depth = 0
    in_str = False
    i = start
    while i < len(text):
        ch = text[i]
        if ch == '"' and text[i - 1] != "\\":
            in_str = not in_str
        elif not in_str:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    return i + 1
        i += 1
    raise AssertionError("unbalanced parenthesis in the Julia shim")


def _c_kind(ctype):
    t = ctype.strip()
    if "*" in t:
        return "ptr"
    t = re.sub(r"\bconst\b", "", t).split()
    base = t[0] if t else ""
    return {"int64_t": "i64", "double": "f64", "int": "i32", "int32_t": "i32", "uint64_t": "i64"}.get(base, base)


def _julia_kind(jtype):
    t = jtype.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Int64": "i64", "Float64": "f64", "Cint": "i32", "Int32": "i32", "UInt64": "i64"}.get(t, t)


def header_prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|const char\s*\*)\s+(gbm_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = " ".join(args.split())
        kinds = []
        if args and args != "void":
            for a in _split_top_level(args):
                a = re.sub(r"/\*.*?\*/", "", a)
                # drop the parameter name: everything up to the last identifier
                m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)?$", a.strip())
                ctype = m.group(1) if m.group(2) and m.group(1).strip() else a
                kinds.append(_c_kind(ctype))
        protos[name] = ("ptr" if "*" in ret else "i32", kinds)
    return protos


@pytest.mark.parametrize("shim", ["genomicbreedingmodels.jl_b200/julia/GenomicBreedingModelsB200.jl", "INTEGRATION.md"])
def test_julia_shim_ccalls_match_the_header(shim):
    """Every `ccall((:gbm_x, LIBGBM), Ret, (ArgTypes...), args...)` of the Julia shim (never executed here) binds a
    symbol the header declares, with the header's return kind, the same number of parameters, the same kind (pointer /
    Int64 / Cint / Float64) in every position, and as many actual arguments as declared types."""
    text = open(os.path.join(ROOT, shim)).read()
    protos = header_prototypes()
    assert len(protos) >= 40
    seen = set()
    for m in re.finditer(r"ccall\(\(:(gbm_[a-z0-9_]+),\s*LIBGBM\)\s*,", text):
        name = m.group(1)
        assert name in protos, f"{name}: not declared in include/gbm_b200.h"
        call_open = text.index("(", m.start())
        body = text[call_open + 1:_balanced(text, call_open) - 1]
        parts = _split_top_level(re.sub(r"#[^\n]*", "", body))
        ret, argtypes, actual = parts[1], parts[2], parts[3:]
        assert argtypes.startswith("(") and argtypes.endswith(")"), (name, argtypes)
        jkinds = [_julia_kind(t) for t in _split_top_level(argtypes[1:-1])]
        want_ret, ckinds = protos[name]
        assert _julia_kind(ret) == want_ret, (name, ret)
        assert jkinds == ckinds, f"{name}: Julia {jkinds} vs header {ckinds}"
        assert len(actual) == len(ckinds), f"{name}: {len(actual)} arguments for {len(ckinds)} parameters"
        seen.add(name)
    if shim.endswith(".md"):  # the binding shown to maintainers: the same check on its snippets
        assert {"gbm_init", "gbm_colstats", "gbm_grm", "gbm_kstd_pc1", "gbm_scan", "gbm_sharded_gwas"} <= seen
        return
    # the hot path's entry points are all bound
    for must in ("gbm_init", "gbm_matrix_upload_compact", "gbm_matrix_upload_indexed", "gbm_colstats", "gbm_grm",
                 "gbm_kstd_pc1", "gbm_scan", "gbm_group_create_local", "gbm_sharded_upload", "gbm_sharded_gwas",
                 "gbm_lmm_plan_create", "gbm_lmm_plan_run", "gbm_transform1_screen", "gbm_transform2_screen"):
        assert must in seen, must


def test_julia_shim_timing_struct_matches_the_header():
    """`GwasTiming` (passed by reference to gbm_sharded_gwas) has gbm_gwas_timing's fields in order and width."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct gbm_gwas_timing \{(.*?)\} gbm_gwas_timing;", src, flags=re.S).group(1)
    c_fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        ctype, names = decl.split(" ", 1)
        c_fields += [(n.strip(), _c_kind(ctype)) for n in names.split(",")]
    text = open(os.path.join(ROOT, "genomicbreedingmodels.jl_b200/julia/GenomicBreedingModelsB200.jl")).read()
    jbody = re.search(r"struct GwasTiming[^\n]*\n(.*?)\nend", text, flags=re.S).group(1)
    j_fields = [(n, _julia_kind(t)) for n, t in re.findall(r"([a-z_0-9]+)::([A-Za-z0-9]+)", jbody)]
    assert j_fields == c_fields


def test_ctypes_signatures_match_the_header():
    """gbm_b200/_lib.py: restype and the kind (pointer / int64 / int / double) of every argtype equal the header's."""
    import ctypes

    from gbm_b200 import _lib as L

    def kind(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or isinstance(t, type(ctypes.POINTER(ctypes.c_int))) and hasattr(t, "_type_") and not isinstance(t._type_, str):
            return "ptr"
        return {ctypes.c_int64: "i64", ctypes.c_uint64: "i64", ctypes.c_int: "i32", ctypes.c_int32: "i32",
                ctypes.c_double: "f64"}[t]

    protos = header_prototypes()
    assert sorted(protos) == sorted(L.SIGNATURES)
    for name, (res, args) in L.SIGNATURES.items():
        want_ret, ckinds = protos[name]
        assert kind(res) == want_ret, name
        assert [kind(a) for a in args] == ckinds, f"{name}: ctypes {[kind(a) for a in args]} vs header {ckinds}"


# ---- host half of the Lanczos PC1 solver (csrc/lanczos.cu: tridiag_top), no GPU needed ------------------------------
def _tridiag_top(alpha, beta):
    import ctypes

    from gbm_b200 import _lib as L

    lib = L.load()
    m = len(alpha)
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    b = np.ascontiguousarray(beta if m > 1 else np.zeros(1), dtype=np.float64)
    s = np.empty(m)
    theta = ctypes.c_double()
    L.check(lib.gbm_tridiag_top(L.ptr(a), L.ptr(b), m, ctypes.byref(theta), L.ptr(s)))
    return theta.value, s


@pytest.mark.parametrize("m", [1, 2, 3, 17, 120, 400])
def test_tridiagonal_top_eigenpair_matches_lapack(m):
    """Sturm bisection + inverse iteration against numpy's eigh on the dense tridiagonal: random matrices, a Lanczos-like
    one (a large leading entry coupled weakly to a clustered bulk, as for the standardised GRM) and one with a nearly
    double top eigenvalue (the vector is then only defined up to the gap; the residual still has to be tiny)."""
    rng = np.random.default_rng(m)
    cases = []
    cases.append((rng.normal(size=m), np.abs(rng.normal(size=max(m - 1, 0)))))
    a = 1.0 + 0.01 * rng.normal(size=m)
    a[0] = 300.0
    b = 0.3 * np.abs(rng.normal(size=max(m - 1, 0))) + 0.05
    cases.append((a, b))
    if m >= 3:
        a2 = np.concatenate([[5.0, 5.0 + 1e-9], rng.uniform(0, 1, size=m - 2)])
        b2 = np.concatenate([[1e-10], 0.1 * np.abs(rng.normal(size=m - 2))])[: m - 1]
        cases.append((a2, b2))
    for a, b in cases:
        T = np.diag(a) + np.diag(b, 1) + np.diag(b, -1) if m > 1 else np.array([[a[0]]])
        w, V = np.linalg.eigh(T)
        theta, s = _tridiag_top(a, b)
        scale = max(np.abs(w).max(), 1e-300)
        assert abs(theta - w[-1]) <= 4e-15 * scale
        assert abs(np.linalg.norm(s) - 1.0) < 1e-12
        assert np.linalg.norm(T @ s - theta * s) <= 2e-13 * scale     # what the solver's stopping test relies on
        gap = (w[-1] - w[-2]) / scale if m > 1 else 1.0
        if gap > 1e-6:  # a separated top eigenvalue: the vector itself
            v = V[:, -1]
            assert min(np.abs(s - v).max(), np.abs(s + v).max()) <= 1e-12 / gap


def test_tridiagonal_top_argument_errors():
    from gbm_b200 import _lib as L

    lib = L.load()
    s = np.empty(2)
    import ctypes

    theta = ctypes.c_double()
    assert lib.gbm_tridiag_top(None, None, 2, ctypes.byref(theta), L.ptr(s)) == L.GBM_ERR_ARGUMENT
    assert lib.gbm_tridiag_top(L.ptr(np.ones(2)), L.ptr(np.ones(1)), 0, ctypes.byref(theta), L.ptr(s)) == L.GBM_ERR_ARGUMENT
