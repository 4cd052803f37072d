"""BASELINE configs[0] (the reference's own CPU-runnable case): gwasols / gwaslmm on n = 300, l = 10,000 through
the host mirror, wall time per call after a warm-up (GPU box; not a test)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "genomicbreedingmodels.jl_b200"))
import numpy as np
import gbm_b200
from oracle import synth, cbind, gwas_oracle as go
gbm_b200.init(0)
n, p = 300, 10000
for kind, name in ((synth.KIND_CONTINUOUS, "continuous allele frequencies"), (synth.KIND_TETRAPLOID, "tetraploid dosages")):
    A = synth.block(42, n, 0, p, kind)
    y = synth.phenotype(42, n, p, kind)
    g = gbm_b200.Genomes.from_matrix(A)
    ph = gbm_b200.Phenomes.from_matrix(y, entries=g.entries)
    for fn in (gbm_b200.gwasols, gbm_b200.gwaslmm):
        fn(genomes=g, phenomes=ph)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); fit = fn(genomes=g, phenomes=ph); ts.append(time.perf_counter() - t0)
        print(f"{name}: {fn.__name__} n={n} l={p}: {np.median(ts)*1e3:.1f} ms per call ({fit.extras['storage']})", flush=True)
    cbind.use_all_cores()
    t0 = time.perf_counter(); b, prep, pc = go.gwasols(A, g.entries, y[:, None], ph.entries); t1 = time.perf_counter()
    print(f"{name}: NumPy/C oracle gwasols (GRM + PCA + literal pinv loop) {1e3*(t1-t0):.1f} ms on {cbind.num_threads()} threads", flush=True)

# per-call times: are there sporadic stalls?
A = synth.block(42, n, 0, p, synth.KIND_TETRAPLOID)
y = synth.phenotype(42, n, p, synth.KIND_TETRAPLOID)
g = gbm_b200.Genomes.from_matrix(A); ph = gbm_b200.Phenomes.from_matrix(y, entries=g.entries)
for fn in (gbm_b200.gwasols, gbm_b200.gwaslmm, gbm_b200.gwasols, gbm_b200.gwaslmm):
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); fn(genomes=g, phenomes=ph); ts.append(1e3 * (time.perf_counter() - t0))
    print(fn.__name__, " ".join(f"{t:.0f}" for t in ts), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): gbm_b200.gwaslmm(genomes=g, phenomes=ph)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
