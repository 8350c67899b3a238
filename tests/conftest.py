import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree artefacts exist (they are prebuilt and travel with the snapshot;
    this only compiles what is missing)."""
    need = [
        os.path.join(ROOT, "cactus-gfa-tools_b200", "lib", "libg2p.so"),
        os.path.join(ROOT, "build", "libgafgen.so"),
        os.path.join(ROOT, "build", "g2p_hostsim"),
        os.path.join(ROOT, "build", "g2p_simt"),
        os.path.join(ROOT, "build", "g2p_simt_long"),
        os.path.join(ROOT, "build", "g2u_hostsim"),
    ]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", ROOT], stdout=subprocess.DEVNULL)
    if not all(os.path.exists(os.path.join(ROOT, "oracle", "bin", b)) for b in ("gaf2paf_oracle", "gaf2unstable_oracle")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    yield


@pytest.fixture(scope="session")
def g2p():
    import cactus_gfa_tools_b200 as m
    return m


@pytest.fixture(scope="session")
def converter(g2p):
    cv = g2p.Converter(0)
    yield cv
    cv.close()
