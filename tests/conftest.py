import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    """True when libg2p.so can create a context on device 0 (no torch import needed)."""
    try:
        import ctypes
        import cactus_gfa_tools_b200 as m
        h = ctypes.c_void_p()
        rc = m.lib.g2p_create(0, ctypes.byref(h))
        if rc == 0:
            m.lib.g2p_destroy(h)
        return rc == 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not errored) on a box without a CUDA device, unless they were
    asked for explicitly with -m gpu: there a missing device must fail loudly (no CPU fallback)."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run with -m gpu on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)


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
        os.path.join(ROOT, "build", "gaf2paf_stub"),
        os.path.join(ROOT, "build", "g2p_filter_simt"),
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
