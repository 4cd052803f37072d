"""Generates tests/golden/*.npz with the CPU oracle (oracle/gwas_oracle.py).

The reference (GenomicBreedingModels.jl) holds no numeric golden vectors for the GWAS path
and cannot be executed in this image (no Julia), so these fixtures pin the ORACLE: the
oracle is re-checked against them in the CPU suite (a regression guard for the checker)
and the CUDA path is checked against them in the GPU suite.  Parity with the real
reference stays "unpinned" for the rows SURVEY.md 8c lists.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import gwas_oracle as go, synth  # noqa: E402

CASES = {
    "tetraploid_n48_p96": dict(seed=42, n=48, p=96, kind=synth.KIND_TETRAPLOID),
    "diploid_n64_p256": dict(seed=7, n=64, p=256, kind=synth.KIND_DIPLOID),
    "continuous_n33_p50": dict(seed=3, n=33, p=50, kind=synth.KIND_CONTINUOUS),
}


def make(case):
    seed, n, p, kind = case["seed"], case["n"], case["p"], case["kind"]
    A = synth.block(seed, n, 0, p, kind)
    y = synth.phenotype(seed, n, p, kind, n_causal=4)
    ent = [f"entry_{i + 1}" for i in range(n)]
    out = dict(A=A, y=y, seed=seed, kind=kind)
    for grm_type, tag in (("simple", "s"), ("ploidy-aware", "p")):
        b_lit, prep, pc = go.gwasols(A, ent, y[:, None], ent, GRM_type=grm_type)
        z, _, _ = go.gwaslmm(A, ent, y[:, None], ent, GRM_type=grm_type)
        raw = A[:, prep.idx_cols - 1]
        cf = go.scan_closed_form(raw, prep.y, pc)
        out[f"idx_cols_{tag}"] = prep.idx_cols
        out[f"K_{tag}"] = prep.K
        out[f"pc1_{tag}"] = pc
        out[f"b_ols_literal_{tag}"] = b_lit
        out[f"b_ols_{tag}"] = cf["stat_ols"]
        out[f"z_lmm_{tag}"] = z
        out[f"beta_{tag}"] = cf["beta"]
        out[f"se_ols_{tag}"] = cf["se_ols"]
        out[f"nlp_t_{tag}"] = go.neglog10_sf_t(cf["stat_ols"], n - 1)
        out[f"nlp_z_{tag}"] = go.neglog10_sf_normal(z)
        if prep.ploidy is not None:
            out["ploidy"] = prep.ploidy
    out["ys"] = prep.y
    out["col_mean"] = prep.col_mean
    out["col_sd"] = prep.col_sd
    out["grm_simple"] = go.grm_simple(A)
    out["grm_simple_uncentred"] = go.grm_simple(A, center=False)
    out["grm_ploidy4"] = go.grm_ploidy_aware(A, 4)
    return out


if __name__ == "__main__":
    for name, case in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **make(case))
        print("wrote", name)
