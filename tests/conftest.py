import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full-size parity checks")


@pytest.fixture(scope="session")
def oracle_c():
    from oracle import oracle_c as oc
    oc.load()
    return oc


@pytest.fixture(scope="session")
def oracle_np():
    from oracle import oracle_np as on
    return on


@pytest.fixture(scope="session")
def engine():
    """One IdleEngine on cuda:0 with room for host windows used by the GPU tests."""
    import gpu_pruner_b200 as g
    eng = g.IdleEngine(device=0, max_pods=20000, max_gpus=8, max_samples=4096, power_plane=True)
    yield eng
    eng.close()
