"""Generates tests/golden/transform_*.npz with the CPU oracle (oracle/transform_oracle.py): the
l (transform1) and l^2 (transform2) screen effects of every named endofunction on a small seeded
problem, including two complementary alleles (addnorm of them is exactly constant: the
rank-deficient branch of Julia's `\\`) and a low-variance locus (skipped by the sigma^2 filter).

The reference holds no numeric vector for this path and Julia is not available here, so the
fixture pins the ORACLE (CPU suite) and the CUDA path is checked against it (GPU suite).

    python tests/golden/make_golden_transform.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import synth, transform_oracle as to  # noqa: E402


def problem(seed, n, l, kind):
    A = synth.block(seed, n, 0, l, kind)
    A[:, 3] = 1.0 - A[:, 2]  # complementary allele of locus 3 (1-based 3 and 4)
    A[:, 5] = 0.0
    A[0, 5] = 0.5  # var = 0.25 / n < 0.01: skipped
    rng = np.random.default_rng(seed)
    b = np.zeros(l)
    b[[1, 2, 7]] = (1.5, -2.0, 1.0)
    y = 10.0 + A @ b + 2.0 * A[:, 1] * A[:, 7] + rng.normal(scale=0.5, size=n)
    return np.asfortranarray(A), y


CASES = {"transform_tetraploid_n40_l12": dict(seed=11, n=40, l=12, kind=synth.KIND_TETRAPLOID),
         "transform_continuous_n57_l9": dict(seed=5, n=57, l=9, kind=synth.KIND_CONTINUOUS)}

if __name__ == "__main__":
    for name, c in CASES.items():
        A, y = problem(**c)
        out = dict(A=A, y=y)
        for f in to.TRANSFORMATIONS1:
            beta, idx, T = to.transform1(f, A, y, n_new=min(6, c["l"]))
            out[f"beta1_{f.__name__}"] = beta
            out[f"idx1_{f.__name__}"] = idx
            out[f"T1_{f.__name__}"] = T
        for f in to.TRANSFORMATIONS2:
            for comm in (False, True):
                beta, idx, pairs, T = to.transform2(f, A, y, n_new=10, commutative=comm)
                tag = f"{f.__name__}_{int(comm)}"
                out[f"beta2_{tag}"] = beta
                out[f"idx2_{tag}"] = idx
                out[f"T2_{tag}"] = T
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name)
