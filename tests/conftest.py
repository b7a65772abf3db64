import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")
BUNDLED = ["cornell-box", "MaterialsScene", "materialball", "coffee", "bathroom"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_flat_cache = {}
_ref_cache = {}


def have_ref():
    from oracle import ref
    return ref.available()


def ref_scene(name):
    """The scene loaded by the UNMODIFIED reference (oracle/_ref); skips when absent."""
    from oracle import ref
    if not ref.available() or not ref.have_scene(name):
        pytest.skip("oracle/_ref (reference build + staged scenes) not available for %s" % name)
    if name not in _ref_cache:
        _ref_cache[name] = ref.RefScene(name)
    return _ref_cache[name]


def flat_scene(name):
    """FlatScene of a bundled scene: the committed fixture for cornell-box, otherwise the
    product flattener run on the reference's Scene (oracle/_ref)."""
    from raytracingrenderer_b200 import abi
    if name not in _flat_cache:
        if name == "cornell-box":
            _flat_cache[name] = abi.FlatScene.load(os.path.join(GOLDEN, "cornell-box.rtbs"))
        else:
            rs = ref_scene(name)
            cache = os.path.join("/tmp", "rtb_test_cache")
            os.makedirs(cache, exist_ok=True)
            _flat_cache[name] = rs.flatten(os.path.join(cache, name + ".rtbs"))
    return _flat_cache[name]


def synthetic_scene(seed=3, n_tris=64, width=96, height=64, with_env=True):
    """Small procedural flat scene with every BSDF class, an area light and a tiny env map,
    its reference-style BVH built by tests/refbvh.py (independent numpy restatement of
    Geometry.h:325-392).  Works without the reference, so GPU-vs-oracle tests always run."""
    import refbvh
    return refbvh.random_scene(seed, n_tris, width, height, with_env)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import port
    port.build()
    return port


@pytest.fixture(scope="session")
def rtb():
    import raytracingrenderer_b200 as m
    return m


def rel_err(a, b, floor=1e-6):
    """max |a-b| / max(|b| row-wise inf-norm, floor) for vectors [n, k] or scalars [n]."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.ndim == 1:
        a, b = a[:, None], b[:, None]
    scale = np.maximum(np.abs(b).max(axis=1), floor)
    return (np.abs(a - b).max(axis=1) / scale).max() if len(a) else 0.0
