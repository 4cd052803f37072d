import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "genomicbreedingmodels.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs at least 2 B200s on the box (gpurun --gpus N); only generated "
                                       "when they are visible, always carries the gpu marker too")


@pytest.fixture(scope="session")
def gbm():
    """The product package with the CUDA library initialised on cuda:0.  Fails loudly
    (no skip, no fallback) when the library or the GPU is missing."""
    import gbm_b200

    gbm_b200.init(0)
    return gbm_b200
