import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: larger CPU case")


_DATA_CACHE = {}


@pytest.fixture(scope="session")
def golden_data():
    """Session cache of reference-recipe data sets keyed by (ntrain, nblocks, local_dist)."""
    from oracle.synthetic import SampledData, golden_run
    from oracle.blocking import grid_centers

    def get(ntrain, nblocks, local_dist, seed=0):
        key = (ntrain, seed)
        if key not in _DATA_CACHE:
            _DATA_CACHE[key] = golden_run(ntrain, nblocks, local_dist, seed=seed)
        sd = _DATA_CACHE[key]
        sd.set_centers(grid_centers(nblocks))
        return sd
    return get
