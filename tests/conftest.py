import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cref():
    """The C oracle (oracle/cpu_ref.c), built on demand.  Test infrastructure only."""
    import cref as R

    R.build()
    return R


@pytest.fixture(scope="session")
def zk():
    """The product: zk_b200 over libzk_b200.so (built on demand; must never fall back to the oracle)."""
    import subprocess

    so = os.path.join(ROOT, "zk_b200", "libzk_b200.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", ROOT, "-j8"], check=True, capture_output=True)
    import zk_b200

    return zk_b200


@pytest.fixture(scope="session")
def ctx(zk):
    return zk.Context.default()


def hx(v):
    return int(v, 16)
