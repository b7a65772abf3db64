#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle/): builds the CPU oracle from the UNMODIFIED reference.

Recipe, outputs only under oracle/_ref/ (git-ignored, travels to the
GPU box with gpurun):

  oracle/_ref/shadow/        symlinks to /root/reference/RTBase/*.h + the two stub headers
  oracle/_ref/shadow_d0/     same, but Renderer.h is a build-time patched COPY whose
                             `const int MAX_DEPTH = 4;` (Renderer.h:20) reads RT_MAX_DEPTH
                             (soup config, SURVEY 8d cfg 5)
  oracle/_ref/librtref.so    oracle/ref_driver.cpp, g++ -O2 -ffp-contract=off  (parity oracle)
  oracle/_ref/librtref_fast.so   same, -O3 -march=x86-64-v3 -ffast-math (speed only; mirrors the
                             reference's MSVC /fp:fast /arch:AVX, RTBase/RTBase.vcxproj:123-124)
  oracle/_ref/librtref_d0.so     MAX_DEPTH = 0 variant of the parity build
  oracle/_ref/librtref_mis.so    parity build whose pathTrace calls computeDirectMIS (Renderer.h:474-557,
                                 shipped but switched off) instead of computeDirect: a build-time patched
                                 COPY of Renderer.h in the git-ignored shadow dir, like _d0
  oracle/_ref/dropin_main    oracle/dropin_main.cpp: reference host program + the product's
                             host/Renderer.h, linked to librtb200.so (drop-in test)
  scenes/_staged/<name>/     filtered copies of the bundled scene ASSETS (instances whose mesh is
                             listed in /root/reference/.MISSING_LARGE_BLOBS removed from
                             scene.json) + BSDF-override variants; git-ignored input data read by
                             both the oracle and the product's own loader

/root/reference is only read.  On the GPU box (no /root/reference) this script is a no-op
and the prebuilt files are used.
"""
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("RTB_REFERENCE_DIR", "/root/reference")
RTBASE = os.path.join(REF, "RTBase")
OUT = os.path.join(HERE, "_ref")
# scene ASSETS (input data, not code) are staged outside oracle/: the product loader reads them too
STAGED = os.path.join(ROOT, "scenes", "_staged")

REF_HEADERS = [
    "Core.h", "Sampling.h", "Geometry.h", "Imaging.h", "Materials.h", "Lights.h", "Scene.h",
    "Renderer.h", "SceneLoader.h", "GEMLoader.h", "stb_image.h", "stb_image_write.h",
]

# Per-instance overrides used for "materialball exercising every BSDF" (SURVEY 8d cfg 2).
MATERIALBALL_VARIANTS = OrderedDict([
    ("diffuse", {"bsdf": "diffuse"}),
    ("conductor", {"bsdf": "conductor", "eta": "0.18 0.42 1.37", "k": "3.42 2.35 1.77", "roughness": "0.02"}),
    ("glass", {"bsdf": "glass", "intIOR": "1.5", "extIOR": "1.0"}),
    ("dielectric", {"bsdf": "dielectric", "intIOR": "1.5", "extIOR": "1.0", "roughness": "0.1"}),
    ("orennayar", {"bsdf": "orennayar", "alpha": "0.5"}),
    ("plastic", {"bsdf": "plastic", "intIOR": "1.5", "extIOR": "1.0", "roughness": "0.1"}),
    ("layered", {"bsdf": "plastic", "intIOR": "1.5", "extIOR": "1.0", "roughness": "0.1",
                 "coatingThickness": "1", "coatingSigmaA": "0.1 0.1 0.1"}),
    ("mirror", {"bsdf": "mirror"}),
])


def have_reference():
    return os.path.isfile(os.path.join(RTBASE, "Renderer.h"))


def make_shadow(name, patch_depth=False, patch_mis=False):
    d = os.path.join(OUT, name)
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.makedirs(os.path.join(d, "OpenImageDenoise"))
    for h in REF_HEADERS:
        src = os.path.join(RTBASE, h)
        dst = os.path.join(d, h)
        if (patch_depth or patch_mis) and h == "Renderer.h":
            text = open(src, encoding="utf-8-sig").read()
            if patch_depth:
                needle = "const int MAX_DEPTH = 4;"
                assert text.count(needle) == 1, "Renderer.h:20 changed"
                text = text.replace(needle, "const int MAX_DEPTH = RT_MAX_DEPTH;")
            if patch_mis:
                # pathTrace calls the estimator the reference ships switched off (Renderer.h:346)
                needle = "Colour direct = pathThroughput * computeDirect(shadingData, sampler);"
                assert text.count(needle) == 1, "Renderer.h:346 changed"
                text = text.replace(needle, "Colour direct = pathThroughput * computeDirectMIS(shadingData, sampler);")
            open(dst, "w", encoding="utf-8").write(text)
        else:
            os.symlink(src, dst)
    shutil.copy(os.path.join(HERE, "stubs", "GamesEngineeringBase.h"), os.path.join(d, "GamesEngineeringBase.h"))
    shutil.copy(os.path.join(HERE, "stubs", "OpenImageDenoise", "oidn.hpp"),
                os.path.join(d, "OpenImageDenoise", "oidn.hpp"))
    return d


def make_dropin_shadow():
    """Like make_shadow, but Renderer.h is the PRODUCT's drop-in header."""
    d = make_shadow("shadow_dropin")
    os.unlink(os.path.join(d, "Renderer.h"))
    os.symlink(os.path.join(ROOT, "raytracingrenderer_b200", "host", "Renderer.h"), os.path.join(d, "Renderer.h"))
    return d


def compile_dropin(shadow):
    """oracle/dropin_main.cpp: reference host program + product RayTracer, linked to librtb200.so."""
    lib_dir = os.path.join(ROOT, "raytracingrenderer_b200")
    if not os.path.isfile(os.path.join(lib_dir, "librtb200.so")):
        return None
    out = os.path.join(OUT, "dropin_main")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-I", shadow,
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "raytracingrenderer_b200", "host"),
                           os.path.join(HERE, "dropin_main.cpp"), "-o", out, "-L", lib_dir, "-lrtb200",
                           "-Wl,-rpath,$ORIGIN/../../raytracingrenderer_b200", "-lpthread"])
    return out


def compile_driver(shadow, out_name, flags):
    out = os.path.join(OUT, out_name)
    cmd = ["g++", "-std=c++17", "-shared", "-fPIC", "-w"] + flags + [
        "-I", shadow, "-I", os.path.join(ROOT, "raytracingrenderer_b200", "host"),
        os.path.join(HERE, "ref_driver.cpp"), "-o", out, "-lpthread",
    ]
    subprocess.check_call(cmd)
    return out


def _filter_scene_json(src_dir, dst_dir, edit=None, prefix=""):
    """Copy scene.json dropping instances whose mesh file is absent (the reference loader
    exit(0)s on them, GEMLoader.h:349-354); `edit(scene)` may change it further."""
    with open(os.path.join(src_dir, "scene.json")) as f:
        scene = json.load(f, object_pairs_hook=OrderedDict)
    dropped = []
    for key, val in scene.items():
        if isinstance(val, list):
            keep = []
            for inst in val:
                if os.path.isfile(os.path.join(src_dir, inst.get("filename", ""))):
                    keep.append(inst)
                else:
                    dropped.append(inst.get("filename"))
            scene[key] = keep
    if prefix:
        for key, val in scene.items():
            if isinstance(val, list):
                for inst in val:
                    inst["filename"] = prefix + inst["filename"]
                    if "reflectance" in inst:
                        inst["reflectance"] = prefix + inst["reflectance"]
        if "envmap" in scene:
            scene["envmap"] = prefix + scene["envmap"]
    if edit:
        edit(scene)
    os.makedirs(dst_dir, exist_ok=True)
    with open(os.path.join(dst_dir, "scene.json"), "w") as f:
        json.dump(scene, f, indent=1)
    return dropped


def stage_scenes():
    scenes_out = STAGED
    os.makedirs(scenes_out, exist_ok=True)
    report = {}
    for name in ["cornell-box", "MaterialsScene", "materialball", "coffee", "bathroom"]:
        src = os.path.join(RTBASE, name)
        dst = os.path.join(scenes_out, name)
        os.makedirs(dst, exist_ok=True)
        for fn in os.listdir(src):
            if fn == "scene.json":
                continue
            s, d = os.path.join(src, fn), os.path.join(dst, fn)
            if not os.path.isfile(d) or os.path.getsize(d) != os.path.getsize(s):
                shutil.copyfile(s, d)
        edit = None
        if name == "MaterialsScene":
            # BASELINE.json cfg 3: lit by RTBase/1.hdr (the scene's own env is a missing blob)
            shutil.copyfile(os.path.join(RTBASE, "1.hdr"), os.path.join(dst, "1.hdr"))

            def edit(scene):
                scene["envmap"] = "1.hdr"
        report[name] = _filter_scene_json(src, dst, edit)
    # MaterialsScene lit by a real lat-long map
    def env_edit(scene):
        scene["envmap"] = "../materialball/envmap.hdr"
    _filter_scene_json(os.path.join(RTBASE, "MaterialsScene"), os.path.join(scenes_out, "MaterialsScene_env"),
                       env_edit, prefix="../MaterialsScene/")
    # cornell-box at 256x256 (64 tiles): small enough for the reference's adaptiveRender in a test
    def small_edit(scene):
        scene["width"], scene["height"] = "256", "256"
    _filter_scene_json(os.path.join(RTBASE, "cornell-box"), os.path.join(scenes_out, "cornell-box_256"),
                       small_edit, prefix="../cornell-box/")
    # materialball with Mesh001's BSDF overridden
    for vname, props in MATERIALBALL_VARIANTS.items():
        def edit(scene, props=props):
            for key, val in scene.items():
                if isinstance(val, list):
                    for inst in val:
                        if inst["filename"].endswith("Mesh001.gem"):
                            for k in ["bsdf", "roughness", "intIOR", "extIOR", "eta", "k", "alpha",
                                      "coatingThickness", "coatingSigmaA"]:
                                inst.pop(k, None)
                            inst.update(props)
        _filter_scene_json(os.path.join(RTBASE, "materialball"), os.path.join(scenes_out, "materialball_" + vname),
                           edit, prefix="../materialball/")
    with open(os.path.join(STAGED, "staged.json"), "w") as f:
        json.dump({"dropped_instances": report}, f, indent=1)
    return report


def build(force=False):
    if not have_reference():
        return False
    os.makedirs(OUT, exist_ok=True)
    stamp = os.path.join(OUT, "librtref.so")
    srcs = [os.path.join(HERE, "ref_driver.cpp"), os.path.join(HERE, "stubs", "GamesEngineeringBase.h"),
            os.path.join(ROOT, "raytracingrenderer_b200", "host", "rtb_flatten.hpp"),
            os.path.join(ROOT, "include", "rtb.h"), os.path.abspath(__file__)]
    fresh = (not force and os.path.isfile(stamp) and os.path.isfile(os.path.join(OUT, "librtref_d0.so"))
             and os.path.isfile(os.path.join(OUT, "librtref_mis.so"))
             and os.path.isfile(os.path.join(OUT, "librtref_fast.so"))
             and all(os.path.getmtime(stamp) >= os.path.getmtime(s) for s in srcs))
    if not fresh:
        shadow = make_shadow("shadow")
        shadow_d0 = make_shadow("shadow_d0", patch_depth=True)
        compile_driver(shadow, "librtref_fast.so", ["-O3", "-march=x86-64-v3", "-ffast-math"])
        compile_driver(shadow_d0, "librtref_d0.so", ["-O2", "-ffp-contract=off", "-DRT_MAX_DEPTH=0"])
        compile_driver(make_shadow("shadow_mis", patch_mis=True), "librtref_mis.so", ["-O2", "-ffp-contract=off"])
        compile_driver(shadow, "librtref.so", ["-O2", "-ffp-contract=off"])
    dropin_srcs = [os.path.join(HERE, "dropin_main.cpp"), os.path.join(ROOT, "raytracingrenderer_b200", "host", "Renderer.h"),
                   os.path.join(ROOT, "raytracingrenderer_b200", "host", "rtb_flatten.hpp"), os.path.join(ROOT, "include", "rtb.h")]
    dropin = os.path.join(OUT, "dropin_main")
    if not os.path.isfile(dropin) or any(os.path.getmtime(dropin) < os.path.getmtime(x) for x in dropin_srcs):
        compile_dropin(make_dropin_shadow())
    if not os.path.isfile(os.path.join(STAGED, "staged.json")) or not fresh:
        stage_scenes()
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "no reference at %s: nothing to build" % REF)
