"""Profiling driver for the pairwise transformation screen (run under ncu on the GPU box; not part
of the tests or the bench): one transform2(mult) and one transform2(raise) screen."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import numpy as np

import gbm_b200
from gbm_b200 import _lib, transform as tr

gbm_b200.init(0)
n, l = 10000, int(os.environ.get("PROF_L", 4096))
dm = gbm_b200.DeviceMatrix.generate(42, n, l, 0)
y = np.random.default_rng(0).normal(size=n)
for f in (tr.mult, tr.addnorm, tr.raise_):
    tr.transform2_screen(dm, y, f, 1000)
    print(f.__name__, _lib.last_timing(), flush=True)
_, idx = tr.transform1_screen(dm, y, tr.log10epsdivlog10eps, 1000)
print("transform1", _lib.last_timing(), flush=True)
