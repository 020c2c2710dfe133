"""Shared fixtures.  Tests marked `gpu` need a B200 and call libtipe_rt.so through its C ABI;
everything else runs on CPU (oracle vs golden vectors, host logic, ABI surface)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own code; present in the build container and (prebuilt) on the GPU box."""
    from oracle.oracle_py import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libtipe_ref.so not available")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, f"golden_{name}.npz")) for name in ("leaves", "primary", "render_libc", "sky")}


@pytest.fixture(scope="session")
def rt():
    """The product: libtipe_rt.so on cuda:0 (fails loudly when the extension or the GPU is missing)."""
    from tipe_raytracer_b200 import Renderer
    r = Renderer()
    yield r
    r.close()
