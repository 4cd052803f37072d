"""Small end-to-end exercise of every kernel, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck python tests/sanitize_case.py

(ragged shapes on purpose: n and p not multiples of the tile sizes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))

import numpy as np

import gbm_b200
from oracle import gwas_oracle as go, synth

gbm_b200.init(0)
n, p = 301, 523
A = synth.block(5, n, 0, p, synth.KIND_TETRAPLOID)
y = synth.phenotype(5, n, p, synth.KIND_TETRAPLOID)
g = gbm_b200.Genomes.from_matrix(A)
ph = gbm_b200.Phenomes.from_matrix(y, entries=g.entries)
f1 = gbm_b200.gwasols(genomes=g, phenomes=ph, GRM_type="ploidy-aware")
f2 = gbm_b200.gwaslmm(genomes=g, phenomes=ph, GRM_type="simple")
f3 = gbm_b200.gwasreml(genomes=g, phenomes=ph, GRM_type="simple")
b_ref, prep, _ = go.gwasols(A, g.entries, y[:, None], ph.entries, GRM_type="ploidy-aware")
assert np.allclose(f1.b_hat, b_ref, rtol=1e-7, atol=1e-8)
dm = gbm_b200.DeviceMatrix.generate(3, 257, 100, synth.KIND_CONTINUOUS, 7)
rng = np.random.default_rng(0)
for T, k in ((1, 0), (3, 1), (5, 2), (13, 1), (20, 1)):
    dm.scan(rng.normal(size=(257, T)), rng.normal(size=(257, k)) if k else None, model=1)
gbm_b200.scan_host(A, y, None, model=0)
dm.free()
sub = gbm_b200.gwasols(genomes=g, phenomes=ph, idx_loci_alleles=np.arange(2, 400, 3))
print("sanitize_case ok", f1.b_hat[:2], f2.b_hat[:2], f3.b_hat[:2], sub.b_hat[:1])
