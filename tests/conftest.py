import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def parity_sd():
    """Seeded parity weights (numpy PCG64) -- the same numbers the golden fixtures were made with."""
    import numpy as np
    from oracle import weights
    g = np.load(os.path.join(GOLDEN, "demo_clip.npz"))
    sd = weights.make_state_dict("parity", int(g["parity_seed"]))
    for k in [k for k in g.files if k.startswith("cks/")]:
        t = sd[k[4:]].double()
        got = np.array([t.sum().item(), t.abs().sum().item()])
        assert np.allclose(got, g[k], rtol=1e-12, atol=1e-9), f"regenerated weights differ from the fixture's: {k}"
    return sd
