import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cfg():
    from larvio_b200.config import Config
    return Config.load(os.path.join(ROOT, "configs", "euroc_mono.yaml"), max_features_in_one_grid=0, sw_size=12)


@pytest.fixture(scope="session")
def seqs(cfg):
    from larvio_b200 import synth
    return [synth.make_sequence(cfg.raw, s, 14) for s in range(2)]


@pytest.fixture(scope="session")
def lib_built():
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "larvio_b200", "lib", "liblarvio_b200.so")
    if not os.path.exists(lib):
        g.build()
    return lib
