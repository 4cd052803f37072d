"""Thin object layer over the C ABI: a device-resident genotype matrix and the kernels
that run on it.  All arithmetic happens in libgbm_b200.so on the GPU."""
from __future__ import annotations

from ctypes import byref, c_double, c_int, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import check, ptr


def _f64(a, order="F"):
    return np.require(a, dtype=np.float64, requirements=[order.upper() + "_CONTIGUOUS", "ALIGNED"])


class DeviceMatrix:
    """n x p column-major Float64 allele-frequency slab resident in HBM (``gbm_matrix``).

    Replaces the host copy ``G::Matrix{Float64} = allele_frequencies[rows, cols]`` of
    ``extractxyetc`` (/root/reference/src/prediction.jl:129)."""

    def __init__(self, handle: c_void_p, n: int, p: int, keepalive=None):
        self._h = handle
        self.n, self.p = int(n), int(p)
        self._keepalive = keepalive
        self.packed = False

    # ---- constructors ----------------------------------------------------------------
    @classmethod
    def upload(cls, A, rows=None, cols=None) -> "DeviceMatrix":
        """From a host array (NumPy, any layout; copied to column-major if needed) or a CUDA
        torch tensor.  ``rows`` / ``cols`` are optional 1-based index vectors."""
        lib = _lib.lib()
        if isinstance(A, np.ndarray):
            if A.ndim != 2:
                raise _lib.ArgumentError("allele frequencies must be a matrix")
            A = _f64(A)
            n0, p0 = A.shape
            lda = n0
        else:  # torch tensor, column-major view expected: shape (p, n) contiguous == n x p col-major
            raise TypeError("upload() takes a NumPy array; use wrap_device() for device memory")
        h = c_void_p()
        if rows is None and cols is None:
            check(lib.gbm_matrix_upload(ptr(A), n0, p0, lda, byref(h)))
            return cls(h, n0, p0)
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        c = None if cols is None else np.ascontiguousarray(cols, dtype=np.int64)
        n = n0 if r is None else r.size
        p = p0 if c is None else c.size
        check(lib.gbm_matrix_upload_indexed(ptr(A), n0, p0, lda, ptr(r), n, ptr(c), p, byref(h)))
        return cls(h, n, p)

    @classmethod
    def upload_compact(cls, A) -> "DeviceMatrix":
        """``gbm_matrix_upload_compact``: the host cores pack the matrix to one-byte dosage codes on the
        way to the device when every element is an exact dosage level (``.packed`` is then True: 1/8 of the
        bytes cross PCIe, the Float64 matrix never exists in HBM); otherwise a Float64 upload."""
        lib = _lib.lib()
        if isinstance(A, np.ndarray):
            if A.ndim != 2:
                raise _lib.ArgumentError("allele frequencies must be a matrix")
            A = _f64(A)
            n, p = A.shape
        else:  # torch CPU tensor of shape (p, n), contiguous == n x p column-major (e.g. pinned)
            p, n = A.shape
        h = c_void_p()
        packed = c_int()
        check(lib.gbm_matrix_upload_compact(ptr(A), n, p, n, byref(h), byref(packed)))
        m = cls(h, n, p)
        m.packed = bool(packed.value)
        return m

    @classmethod
    def wrap_device(cls, data_ptr: int, n: int, p: int, lda: int, keepalive=None) -> "DeviceMatrix":
        h = c_void_p()
        check(_lib.lib().gbm_matrix_wrap(c_void_p(data_ptr), n, p, lda, byref(h)))
        return cls(h, n, p, keepalive)

    @classmethod
    def generate(cls, seed: int, n: int, p: int, kind: int, col0: int = 0) -> "DeviceMatrix":
        """Synthetic columns col0..col0+p-1 made on the device (oracle/synth.py arithmetic)."""
        h = c_void_p()
        check(_lib.lib().gbm_matrix_generate(seed, n, p, col0, kind, byref(h)))
        return cls(h, n, p)

    @classmethod
    def upload_packed(cls, codes: np.ndarray) -> "DeviceMatrix":
        """From one-byte dosage codes already on the host (n x p, uint8, a = code / 240)."""
        codes = np.require(codes, dtype=np.uint8, requirements=["F_CONTIGUOUS", "ALIGNED"])
        n, p = codes.shape
        h = c_void_p()
        check(_lib.lib().gbm_matrix_upload_packed(ptr(codes), n, p, n, byref(h)))
        m = cls(h, n, p)
        m.packed = True
        return m

    def pack(self):
        """One-byte-per-genotype copy of this matrix (``gbm_matrix_pack``), or None when some
        element is not exactly a dosage code (then keep using the Float64 matrix)."""
        h, bad = c_void_p(), c_int64()
        check(_lib.load().gbm_matrix_pack(self._h, byref(h), byref(bad)))
        if not h.value:
            return None
        m = DeviceMatrix(h, self.n, self.p)
        m.packed = True
        return m

    # ---- housekeeping ----------------------------------------------------------------
    def info(self):
        n, p, lda, d = c_int64(), c_int64(), c_int64(), c_void_p()
        check(_lib.load().gbm_matrix_info(self._h, byref(n), byref(p), byref(lda), byref(d)))
        return {"n": n.value, "p": p.value, "lda": lda.value, "device_ptr": d.value}

    def download(self, j0: int = 0, ncols: int | None = None) -> np.ndarray:
        ncols = self.p - j0 if ncols is None else ncols
        out = np.empty((self.n, ncols), dtype=np.float64, order="F")
        check(_lib.load().gbm_matrix_download(self._h, j0, ncols, ptr(out), self.n))
        return out

    def download_into(self, out: np.ndarray, j0: int = 0):
        """Columns j0 .. j0 + out.shape[0] - 1 into ``out``, a C-contiguous (ncols, n) array (= n x ncols
        column-major), without allocating."""
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape[1] == self.n
        check(_lib.load().gbm_matrix_download(self._h, j0, out.shape[0], ptr(out), self.n))

    def download_cols(self, idx_cols=None, standardise: bool = False) -> np.ndarray:
        """G[:, idx_cols] (1-based), optionally column-standardised on the device
        (/root/reference/src/gwas.jl:114, :129)."""
        idx = None if idx_cols is None else np.ascontiguousarray(idx_cols, dtype=np.int64)
        ncols = self.p if idx is None else idx.size
        out = np.empty((self.n, ncols), dtype=np.float64, order="F")
        check(_lib.load().gbm_matrix_download_cols(self._h, ptr(idx), ncols, int(standardise), ptr(out), self.n))
        return out

    def free(self):
        if self._h is not None:
            check(_lib.load().gbm_matrix_free(self._h))
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- kernels ---------------------------------------------------------------------
    def colstats(self):
        """std(G, dims=1), the fixed-locus filter and the ploidy probe
        (/root/reference/src/gwas.jl:112-113, :119).  Returns a dict with mean, sd,
        min_nonzero, keep (bool), idx_cols (1-based int64, ascending), min_nonzero_kept."""
        p = self.p
        mean, sd, mnz = np.empty(p), np.empty(p), np.empty(p)
        keep = np.empty(p, dtype=np.uint8)
        idx = np.empty(p, dtype=np.int64)
        nk, mk = c_int64(), c_double()
        check(_lib.load().gbm_colstats(self._h, ptr(mean), ptr(sd), ptr(mnz), ptr(keep), ptr(idx), byref(nk),
                                       byref(mk)))
        return {"mean": mean, "sd": sd, "min_nonzero": mnz, "keep": keep.astype(bool),
                "idx_cols": idx[: nk.value].copy(), "min_nonzero_kept": mk.value}

    def grm(self, grm_type: int = _lib.GRM_SIMPLE, ploidy: int = 2, flags: int = 0, out=None):
        """Full symmetric GRM (n x n).  ``out`` may be a CUDA torch tensor (n*n float64) to
        keep the result on the device; otherwise a NumPy array is returned.
        Returns (K, tflops)."""
        n = self.n
        K = np.empty((n, n), dtype=np.float64, order="F") if out is None else out
        tf = c_double()
        check(_lib.load().gbm_grm(self._h, grm_type, ploidy, flags, ptr(K), byref(tf)))
        return K, tf.value

    def grm_accumulate(self, dK_ptr: int, centre: bool = True):
        """Marker-shard partial: dK += sum_j (a_j - mu_j)(a_j - mu_j)' (lower triangle).
        Returns (sum_j q_j(1-q_j), tflops)."""
        s, tf = c_double(0.0), c_double()
        check(_lib.load().gbm_grm_accumulate(self._h, int(centre), c_void_p(dK_ptr), byref(s), byref(tf)))
        return s.value, tf.value

    def scan(self, Y, C=None, model: int = _lib.MODEL_OLS, flags: int = 0, want=("beta", "se", "stat", "neglog10p")):
        """Per-marker association scan (/root/reference/src/gwas.jl:239-249, :363-389).
        Y: n or n x T (used as given); C: n x k covariates without intercept (PC1)."""
        Y = np.asarray(Y, dtype=np.float64)
        if Y.shape[0] != self.n:
            raise _lib.ArgumentError("phenotype length does not match the number of entries")
        Y = _f64(Y.reshape(self.n, -1))
        T = Y.shape[1]
        if C is None:
            Cm, k = None, 0
        else:
            Cm = _f64(np.asarray(C, dtype=np.float64).reshape(self.n, -1))
            k = Cm.shape[1]
        p = self.p
        out = {name: np.empty((p, T), dtype=np.float64, order="F") for name in want}
        mean, sd = np.empty(p), np.empty(p)
        keep = np.empty(p, dtype=np.uint8)
        check(_lib.load().gbm_scan(self._h, ptr(Y), T, self.n, ptr(Cm), k, self.n, model, flags,
                                   ptr(out.get("beta")), ptr(out.get("se")), ptr(out.get("stat")),
                                   ptr(out.get("neglog10p")), ptr(mean), ptr(sd), ptr(keep)))
        out.update(mean=mean, sd=sd, keep=keep.astype(bool))
        return out


class ScanPlan:
    """Prepared scan over a resident DeviceMatrix (``gbm_scan_plan``): run() is just the two
    kernel launches.  Outputs may be NumPy arrays (copied out) or CUDA tensors (written in place)."""

    def __init__(self, dm: DeviceMatrix, Y, C=None, model: int = _lib.MODEL_OLS, flags: int = 0):
        Y = np.asarray(Y, dtype=np.float64)
        if Y.shape[0] != dm.n:
            raise _lib.ArgumentError("phenotype length does not match the number of entries")
        Y = _f64(Y.reshape(dm.n, -1))
        Cm = None if C is None else _f64(np.asarray(C, dtype=np.float64).reshape(dm.n, -1))
        self.dm, self.T = dm, Y.shape[1]
        h = c_void_p()
        check(_lib.load().gbm_scan_plan_create(dm._h, ptr(Y), self.T, dm.n, ptr(Cm), 0 if Cm is None else Cm.shape[1],
                                               dm.n, model, flags, byref(h)))
        self._h = h

    def run(self, beta=None, se=None, stat=None, neglog10p=None, mean=None, sd=None, keep=None):
        check(_lib.load().gbm_scan_plan_run(self._h, ptr(beta), ptr(se), ptr(stat), ptr(neglog10p), ptr(mean),
                                            ptr(sd), ptr(keep)))
        return _lib.last_timing()

    def free(self):
        if self._h is not None:
            check(_lib.load().gbm_scan_plan_free(self._h))
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class LmmPlan:
    """GRM-covariance LMM scan (``gbm_lmm_plan``): eigendecomposition of the symmetric GRM,
    rotation of y / covariates, null-model delta; run() rotates a resident DeviceMatrix in
    column blocks (DMMA GEMM) and searches delta per marker.  Model of the reference's
    gwasreml (/root/reference/src/gwas.jl:549-613), see oracle/lmm_oracle.py for the definition."""

    def __init__(self, K, y, C=None):
        n = int(np.asarray(y).shape[0])
        y = np.ascontiguousarray(y, dtype=np.float64)
        Cm = None if C is None else _f64(np.asarray(C, dtype=np.float64).reshape(n, -1))
        if isinstance(K, np.ndarray):
            K = _f64(K)
            if K.shape != (n, n):
                raise _lib.ArgumentError("GRM must be n x n")
        h = c_void_p()
        eig, lam0 = c_double(), c_double()
        check(_lib.lib().gbm_lmm_plan_create(ptr(K), n, ptr(y), ptr(Cm), 0 if Cm is None else Cm.shape[1], n, byref(h),
                                             byref(eig), byref(lam0)))
        self._h, self.n = h, n
        self.eig_ms, self.null_log_delta = eig.value, lam0.value

    def run(self, dm: DeviceMatrix, flags: int = 0):
        p = dm.p
        out = {k: np.empty(p) for k in ("beta", "se", "stat", "neglog10p", "log_delta")}
        tf, sms = c_double(), c_double()
        check(_lib.load().gbm_lmm_plan_run(self._h, dm._h, flags, ptr(out["beta"]), ptr(out["se"]), ptr(out["stat"]),
                                           ptr(out["neglog10p"]), ptr(out["log_delta"]), byref(tf), byref(sms)))
        out["gemm_tflops"], out["search_ms"] = tf.value, sms.value
        out["timing"] = _lib.last_timing()
        return out

    def free(self):
        if self._h is not None:
            check(_lib.load().gbm_lmm_plan_free(self._h))
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def gemm_tn(dA_ptr: int, lda: int, dB_ptr: int, ldb: int, dC_ptr: int, ldc: int, M: int, N: int, K: int) -> float:
    """C = A'B on the DMMA rotation GEMM (device pointers); returns TFLOP/s."""
    tf = c_double()
    check(_lib.lib().gbm_gemm_tn(c_void_p(dA_ptr), lda, c_void_p(dB_ptr), ldb, c_void_p(dC_ptr), ldc, M, N, K, byref(tf)))
    return tf.value


def grm_finalize(dK_ptr: int, n: int, scale: float):
    check(_lib.lib().gbm_grm_finalize(c_void_p(dK_ptr), n, scale))


def kstd_pc1(K, want_kstd: bool = True, want_pc1: bool = True):
    """K column-standardisation (/root/reference/src/gwas.jl:130) and PC1 of the result
    (MultivariateStats PCA, gwas.jl:234).  K: NumPy n x n.  Returns (Kstd | None, pc1, eig_ms)."""
    K = _f64(K)
    n = K.shape[0]
    if K.shape != (n, n):
        raise _lib.ArgumentError("GRM must be square")
    Ks = np.empty((n, n), dtype=np.float64, order="F") if want_kstd else None
    pc = np.empty(n) if want_pc1 else None
    ms = c_double(0.0)
    check(_lib.lib().gbm_kstd_pc1(ptr(K), n, ptr(Ks), ptr(pc), byref(ms)))
    return Ks, pc, ms.value


def kstd_pc1_device(dK_ptr: int, n: int):
    """Same from a DEVICE n x n GRM; only PC1 comes back to the host."""
    pc = np.empty(n)
    ms = c_double()
    check(_lib.lib().gbm_kstd_pc1(c_void_p(dK_ptr), n, None, ptr(pc), byref(ms)))
    return pc, ms.value


def pack_host(A: np.ndarray):
    """Host-side packer (all cores): returns (codes uint8 n x p, n_inexact)."""
    A = _f64(A)
    n, p = A.shape
    out = np.empty((n, p), dtype=np.uint8, order="F")
    bad = c_int64()
    check(_lib.load().gbm_pack_host(ptr(A), n, p, n, ptr(out), n, byref(bad)))
    return out, int(bad.value)


def scan_host(A, Y, C=None, model: int = _lib.MODEL_OLS, flags: int = 0, pack: bool = True):
    """End-to-end scan from a HOST matrix (pinned or pageable): H2D in column blocks
    overlapped with the scan kernel.  pack=True (default): blocks whose elements are all
    dosage codes are packed to one byte per genotype by the host cores before crossing PCIe
    (identical results); pack=False forces plain Float64 copies."""
    if not pack:
        flags |= _lib.SCAN_HOST_NO_PACK
    if isinstance(A, np.ndarray):
        A = _f64(A)
        n, p = A.shape
        a_ptr = ptr(A)
    else:  # torch CPU tensor of shape (p, n), contiguous  == n x p column-major
        p, n = A.shape
        a_ptr = ptr(A)
    Y = _f64(np.asarray(Y, dtype=np.float64).reshape(n, -1))
    T = Y.shape[1]
    Cm = None if C is None else _f64(np.asarray(C, dtype=np.float64).reshape(n, -1))
    k = 0 if Cm is None else Cm.shape[1]
    out = {name: np.empty((p, T), dtype=np.float64, order="F") for name in ("beta", "se", "stat", "neglog10p")}
    mean, sd = np.empty(p), np.empty(p)
    keep = np.empty(p, dtype=np.uint8)
    check(_lib.lib().gbm_scan_host(a_ptr, n, p, n, ptr(Y), T, n, ptr(Cm), k, n, model, flags, ptr(out["beta"]),
                                   ptr(out["se"]), ptr(out["stat"]), ptr(out["neglog10p"]), ptr(mean), ptr(sd),
                                   ptr(keep)))
    out.update(mean=mean, sd=sd, keep=keep.astype(bool))
    return out


def neglog10_sf(stat, dist: str, df: float = 1.0) -> np.ndarray:
    s = np.ascontiguousarray(stat, dtype=np.float64).ravel()
    out = np.empty_like(s)
    d = {"t": 0, "normal": 1}[dist]
    check(_lib.lib().gbm_neglog10_sf(ptr(s), s.size, d, float(df), ptr(out)))
    return out


def measure_copy_bandwidth(nbytes: int = 1 << 30, reps: int = 5) -> float:
    g = c_double()
    check(_lib.lib().gbm_measure_copy_bandwidth(nbytes, reps, byref(g)))
    return g.value
