"""pytest config: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` tests run on the CPU-only build container; `-m gpu` tests run on a
B200 box (no /root/reference there: every input comes from tests/golden/).
"""
import gzip
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

RES = os.path.join(ROOT, "tests", "golden", "resources")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def resource_bytes(name: str) -> bytes:
    """Fixture file copied (gzip) from the reference's resources/ directory."""
    with gzip.open(os.path.join(RES, name + ".gz"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def resources(tmp_path_factory):
    """Directory with the fixtures unpacked (chain loader takes a path)."""
    d = tmp_path_factory.mktemp("resources")
    for fn in os.listdir(RES):
        if fn.endswith(".gz"):
            with open(os.path.join(d, fn[:-3]), "wb") as f:
                f.write(resource_bytes(fn[:-3]))
    return str(d)


@pytest.fixture(scope="session")
def oracle():
    from oracle_ffi import Oracle, build_oracle

    build_oracle()
    return Oracle()
