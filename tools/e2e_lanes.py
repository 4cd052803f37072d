"""Measurement driver for the two lanes of gbm_scan_host (run on the GPU box; not part of the
tests or the bench): host packer alone, copy engine alone, both, plain Float64 copies, and the
host packer's own rate on the same pinned buffer.  Prints one JSON object per line."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import numpy as np
import torch

import gbm_b200
from gbm_b200 import _lib
from oracle import synth

n = int(os.environ.get("LANES_N", 10000))
pe = int(os.environ.get("LANES_P", 100000))
reps = int(os.environ.get("LANES_REPS", 3))
gbm_b200.init(0)
lib = _lib.lib()
host = torch.empty((pe, n), dtype=torch.float64, pin_memory=True)
sub = gbm_b200.DeviceMatrix.generate(42, n, pe, synth.KIND_DIPLOID)
cudart = ctypes.CDLL("libcudart.so.12")
assert cudart.cudaMemcpy(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(sub.info()["device_ptr"]),
                         ctypes.c_size_t(8 * n * pe), ctypes.c_int(2)) == 0
sub.free()
rng = np.random.default_rng(0)
Y = np.asfortranarray(rng.normal(size=(n, 1)))
C = np.asfortranarray(rng.normal(size=(n, 1)))
hout = {k: np.empty(pe) for k in ("beta", "se", "stat", "nlp", "mean", "sd")}
hkeep = np.empty(pe, dtype=np.uint8)
print(json.dumps({"host_threads": len(os.sched_getaffinity(0)), "n": n, "markers": pe,
                  "pack_isa": os.environ.get("GBM_PACK_ISA", "auto")}), flush=True)

# the host packer alone on the pinned buffer
codes = np.empty((n, pe), dtype=np.uint8, order="F")
bad = ctypes.c_int64()
for i in range(reps + 1):
    t0 = time.perf_counter()
    _lib.check(lib.gbm_pack_host(_lib.ptr(host), n, pe, n, _lib.ptr(codes), n, ctypes.byref(bad)))
    dt = time.perf_counter() - t0
print(json.dumps({"what": "gbm_pack_host alone", "GBps_f64_read": 8.0 * n * pe / dt / 1e9, "inexact": bad.value}), flush=True)
del codes


def step(flags):
    _lib.check(lib.gbm_scan_host(_lib.ptr(host), n, pe, n, _lib.ptr(Y), 1, n, _lib.ptr(C), 1, n, 1, flags,
                                 _lib.ptr(hout["beta"]), _lib.ptr(hout["se"]), _lib.ptr(hout["stat"]),
                                 _lib.ptr(hout["nlp"]), _lib.ptr(hout["mean"]), _lib.ptr(hout["sd"]), _lib.ptr(hkeep)))


ref = None
for name, lanes, flags in (("float64 copies, Float64 kernel (NO_PACK)", None, _lib.SCAN_HOST_NO_PACK),
                           ("copy-engine lane only (device pack)", "copy", 0),
                           ("host lane only (host pack)", "host", 0),
                           ("both lanes", "both", 0),
                           ("default (host lane with >= 12 host threads per rank, else copy-engine lane)", None, 0)):
    if lanes:
        os.environ["GBM_SCAN_HOST_LANES"] = lanes
    else:
        os.environ.pop("GBM_SCAN_HOST_LANES", None)
    step(flags)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        step(flags)
        best = min(best, time.perf_counter() - t0)
    tm = _lib.last_timing()
    out = {"what": name, "markers_per_s": pe / best, "GBps_f64_equiv": 8.0 * n * pe / best / 1e9,
           "packed_blocks": tm["packed_blocks"], "host_packed_blocks": tm["host_packed_blocks"],
           "h2d_bytes": tm["h2d_bytes"]}
    if flags == 0:
        if ref is None:
            ref = {k: v.copy() for k, v in hout.items()}
        else:
            out["identical_to_other_lane"] = all(np.array_equal(ref[k], hout[k], equal_nan=True) for k in ref)
    print(json.dumps(out), flush=True)

# does the copy engine slow the packing cores down?  pack the pinned buffer on all cores while plain
# pinned H2D copies of another pinned buffer run back to back
import threading

os.environ.pop("GBM_SCAN_HOST_LANES", None)
other = torch.empty((pe // 4, n), dtype=torch.float64, pin_memory=True)
other.zero_()
dev = torch.empty_like(other, device="cuda")
codes = np.empty((n, pe), dtype=np.uint8, order="F")
res = {}


def packer():
    t0 = time.perf_counter()
    for _ in range(3):
        _lib.check(lib.gbm_pack_host(_lib.ptr(host), n, pe, n, _lib.ptr(codes), n, ctypes.byref(bad)))
    res["pack_GBps"] = 3 * 8.0 * n * pe / (time.perf_counter() - t0) / 1e9


th = threading.Thread(target=packer)
th.start()
copied = 0
t0 = time.perf_counter()
while th.is_alive():
    dev.copy_(other, non_blocking=True)
    torch.cuda.synchronize()
    copied += other.numel() * 8
res["h2d_GBps"] = copied / (time.perf_counter() - t0) / 1e9
th.join()
res["what"] = "gbm_pack_host and plain pinned H2D copies at the same time"
print(json.dumps(res), flush=True)

# ingestion: gbm_matrix_upload / gbm_matrix_upload_compact from pageable and page-locked memory
pag = np.empty((n, pe // 2), order="F")
pag[:] = host.numpy()[: pe // 2].T
pin = host[: pe // 2]
for name, src, fn in (("upload pageable (staged by host workers)", pag, gbm_b200.DeviceMatrix.upload),
                      ("upload_compact pageable (packed by host workers)", pag, gbm_b200.DeviceMatrix.upload_compact),
                      ("upload_compact page-locked", pin, gbm_b200.DeviceMatrix.upload_compact)):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        m = fn(src)
        best = min(best, time.perf_counter() - t0)
        packed = bool(getattr(m, "packed", False))
        m.free()
    print(json.dumps({"what": name, "GBps_f64_equiv": 8.0 * n * (pe // 2) / best / 1e9, "packed": packed}), flush=True)
