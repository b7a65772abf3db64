"""The host-side tree builders of the product (csrc/rtb_accel.hpp, csrc/rtb_cwbvh.hpp) checked on the CPU: the structural
and CONSERVATIVENESS properties the device traversals rely on for returning the reference's hits — every
reference leaf a primitive exactly once, FAST child boxes exactly the union of their leaves' boxes, CW / Q16 quantised boxes
containing the exact boxes with the stated margin, exact leaf boxes carried bit for bit.  tests/tools/accel_check.cpp is
compiled with g++ against the product headers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, synthetic_scene
from raytracingrenderer_b200 import abi

SRC = os.path.join(ROOT, "tests", "tools", "accel_check.cpp")
OUT = os.path.join("/tmp", "rtb_test_cache", "libaccel_check.so")


@pytest.fixture(scope="module")
def checker():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "raytracingrenderer_b200", "csrc", f) for f in ("rtb_accel.hpp", "rtb_cwbvh.hpp")]
    if not os.path.isfile(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), SRC, "-o", OUT, "-lpthread"])
    lib = C.CDLL(OUT)
    lib.accel_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.accel_check_message.restype = C.c_char_p
    return lib


def run(lib, scene):
    nodes = np.ascontiguousarray(scene.ref_nodes, abi.ref_node_dt)
    stats = np.zeros(9, np.float64)
    rc = lib.accel_check(nodes.ctypes.data, len(nodes), scene.n_tris, stats.ctypes.data)
    assert rc == 0, lib.accel_check_message().decode()
    return dict(zip(("fast_nodes", "fast_depth", "wide_nodes", "cw_nodes", "cw_depth", "cw_leaves", "q16_leaves", "cw_inflation", "q16_inflation"), stats))


def test_trees_of_the_committed_and_synthetic_scenes(checker):
    for s in (abi.FlatScene.load(os.path.join(GOLDEN, "cornell-box.rtbs")), synthetic_scene(), synthetic_scene(seed=11, n_tris=700)):
        st = run(checker, s)
        assert st["cw_leaves"] == st["q16_leaves"] == st["fast_nodes"] + 1        # a binary tree over L leaves has L - 1 nodes
        assert st["cw_nodes"] <= st["fast_nodes"] and st["cw_depth"] <= st["fast_depth"]


@pytest.mark.parametrize("name", ["MaterialsScene", "materialball", "coffee", "bathroom"])
def test_trees_of_the_bundled_scenes(checker, name):
    from raytracingrenderer_b200 import host_api
    d = os.path.join(ROOT, "scenes", "_staged", name)
    if not os.path.isfile(os.path.join(d, "scene.json")):
        pytest.skip("staged scene assets missing")
    st = run(checker, host_api.load_scene(d))
    assert st["cw_depth"] + 2 <= 40                      # RTB_CW_STACK
    assert st["fast_depth"] + 2 <= 96                    # RTB_STACK
    assert st["cw_nodes"] < 0.45 * st["fast_nodes"]      # eight-wide: far fewer nodes than the binary tree


@pytest.mark.parametrize("log2n", [15, 19])
def test_soup_trees(checker, log2n):
    """2^19 triangles (> 2^18 leaves) takes the builder's multi-threaded split path (chunked binning, stable partition)."""
    from raytracingrenderer_b200 import host_api
    s, _ = host_api.build_soup(1 << log2n, 64, 36)
    st = run(checker, s)
    assert st["cw_leaves"] == st["fast_nodes"] + 1


def test_a_broken_tree_is_noticed(checker):
    """The checker itself: a reference BVH whose leaf range is out of bounds is rejected by the builder."""
    s = synthetic_scene()
    nodes = np.ascontiguousarray(s.ref_nodes, abi.ref_node_dt).copy()
    leaf = np.flatnonzero(nodes["a"] < 0)[0]
    nodes["b"][leaf] = 7
    stats = np.zeros(9, np.float64)
    assert checker.accel_check(nodes.ctypes.data, len(nodes), s.n_tris, stats.ctypes.data) < 0


def test_skip_links_equal_the_subtree_ends_and_malformed_trees_are_rejected(checker):
    """rtb_accel::buildSkipLinks (one forward pass) against the definition: skip[i] = i + size of i's subtree, computed here
    by recursion over the (a, b) child links; node arrays that are not one pre-order tree are errors, not walks."""
    import sys
    s = synthetic_scene(seed=5, n_tris=900)
    nodes = np.ascontiguousarray(s.ref_nodes, abi.ref_node_dt).copy()
    n = len(nodes)
    want = np.zeros(n, np.uint32)
    sys.setrecursionlimit(10000)

    def size(i):
        if nodes["a"][i] < 0:
            want[i] = i + 1
            return 1
        k = 1 + size(int(nodes["a"][i])) + size(int(nodes["b"][i]))
        want[i] = i + k
        return k

    assert size(0) == n
    checker.accel_skip_links.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    got = np.zeros(n, np.uint32)
    n_leaves = checker.accel_skip_links(nodes.ctypes.data, n, s.n_tris, got.ctypes.data)
    assert n_leaves == int(((nodes["a"] < 0) & (nodes["b"] > 0)).sum())
    assert np.array_equal(got, want)
    interior = np.flatnonzero(nodes["a"] >= 0)
    for what in ("b_points_back", "b_beyond_subtree", "a_not_next", "orphan_tail", "leaf_where_a_subtree_was"):
        bad = nodes.copy()
        i = int(interior[len(interior) // 2])
        if what == "b_points_back":
            bad["b"][i] = i
        elif what == "b_beyond_subtree":
            bad["b"][i] = want[i]
        elif what == "a_not_next":
            bad["a"][i] = i + 2
        elif what == "orphan_tail":
            bad = np.concatenate([bad, bad[-1:]])
        else:
            bad["a"][i], bad["b"][i] = -1, 1          # an interior node turned into a leaf: its old subtree has no parent
        assert checker.accel_skip_links(bad.ctypes.data, len(bad), s.n_tris, np.zeros(len(bad), np.uint32).ctypes.data) < 0, what
        assert b"pre-order" in checker.accel_check_message()


def test_reoptimised_trees_keep_every_property(checker):
    """RTB_TREE_OPT (rtb_accel::FastOptimizer, insertion-based re-optimisation): the optimised tree passes the same checks
    as the builder's — every reference leaf once, child boxes the exact unions, WIDE / CW / Q16 re-encodings conservative —
    and its summed surface area is not larger."""
    from raytracingrenderer_b200 import host_api
    scenes = [synthetic_scene(seed=11, n_tris=700), host_api.build_soup(1 << 13, 64, 36)[0]]
    d = os.path.join(ROOT, "scenes", "_staged", "coffee")
    if os.path.isfile(os.path.join(d, "scene.json")):
        scenes.append(host_api.load_scene(d))
    checker.accel_set_order.argtypes = [C.c_int, C.c_void_p]
    try:
        for passes, fraction, order in ((1, 1.0, 0), (3, 0.05, 3), (0, 1.0, 3), (2, 7.5, 2), (1, -1.0, 4)):
            checker.accel_set_optimise.argtypes = [C.c_int, C.c_double]
            checker.accel_set_optimise(passes, fraction)
            for s in scenes:
                tris = np.ascontiguousarray(s.tri_isect)
                checker.accel_set_order(order, tris.ctypes.data)     # the any-hit child order (RTB_TREE_ORDER) on top
                st = run(checker, s)
                if passes == 0:
                    continue
                sah = np.zeros(2)
                checker.accel_get_sah(C.c_void_p(sah.ctypes.data))
                assert sah[0] > 0 and sah[1] <= sah[0] * 1.0001, sah
                assert st["fast_depth"] + 2 <= 96
    finally:
        checker.accel_set_optimise(0, 1.0)
        checker.accel_set_order(0, None)
