#!/usr/bin/env python3
"""TEST/BENCH INPUT STAGING (oracle/): runs the product flattener on the reference's own
loadScene output (oracle/_ref) and writes flat scenes to scenes/_cache/ (git-ignored; travels
to the GPU box).  bench.py and the examples read those files as INPUT DATA; nothing under
oracle/ is executed at bench time by the product arm.  Replaced by the stand-alone host
loader for scenes it can already read.

  scenes/_cache/<scene>.rtbs               cornell-box, MaterialsScene, MaterialsScene_env,
                                           materialball, coffee
  scenes/_cache/materialball_variants.npz  the material table of every materialball BSDF
                                           override (geometry/textures identical to the base)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
CACHE = os.path.join(ROOT, "scenes", "_cache")
SMALL = ["cornell-box", "MaterialsScene", "MaterialsScene_env", "materialball", "coffee"]


def stage(force=False):
    from oracle import build_ref, ref
    if not ref.available():
        return False
    os.makedirs(CACHE, exist_ok=True)
    stamp = os.path.join(ref.REF_DIR, "librtref.so")
    for name in SMALL:
        out = os.path.join(CACHE, name + ".rtbs")
        if force or not os.path.isfile(out) or os.path.getmtime(out) < os.path.getmtime(stamp):
            ref.RefScene(name).flatten(out)
    vout = os.path.join(CACHE, "materialball_variants.npz")
    if force or not os.path.isfile(vout) or os.path.getmtime(vout) < os.path.getmtime(stamp):
        from raytracingrenderer_b200 import abi
        base = abi.FlatScene.load(os.path.join(CACHE, "materialball.rtbs"))
        tables = {}
        for v in build_ref.MATERIALBALL_VARIANTS:
            f = ref.RefScene("materialball_" + v).flatten("/tmp/_stage_variant.rtbs")
            for k in ("ref_nodes", "tri_isect", "tri_shade", "textures", "texels", "lights"):
                assert getattr(f, k).tobytes() == getattr(base, k).tobytes(), (v, k)
            tables[v] = f.materials
        np.savez(vout, **tables)
    return True


if __name__ == "__main__":
    print("staged" if stage("--force" in sys.argv) else "oracle/_ref not available: nothing staged")
