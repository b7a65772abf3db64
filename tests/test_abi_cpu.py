"""The C-ABI boundary without a GPU: the built library loads, exports every symbol
include/rtb.h declares, its struct layouts match the Python mirror, and it refuses to work
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def header_text():
    return open(os.path.join(ROOT, "include", "rtb.h")).read()


def declared_functions():
    # prototypes at top level: "<type> rtb_xxx(" at the start of a line
    return sorted(set(re.findall(r"^(?:int|void|const char\*)\s+(rtb_\w+)\s*\(", header_text(), re.M)))


def test_header_functions_are_mirrored(rtb):
    assert declared_functions() == sorted(rtb.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(rtb):
    assert os.path.isfile(rtb.LIB_PATH), "librtb200.so not built (python -m raytracingrenderer_b200.build)"
    out = subprocess.check_output(["nm", "-D", "--defined-only", rtb.LIB_PATH]).decode()
    exported = set(re.findall(r"\sT\s+(rtb_\w+)", out))
    missing = [s for s in declared_functions() if s not in exported]
    assert not missing, missing
    L = rtb.lib()  # loads with ctypes, ABI version checked
    for s in declared_functions():
        assert hasattr(L, s)


def test_struct_layouts_match_python_mirror(tmp_path, rtb):
    from raytracingrenderer_b200 import abi
    names = ["rtb_camera", "rtb_ref_node", "rtb_tri_isect", "rtb_tri_shade", "rtb_material", "rtb_texture",
             "rtb_light", "rtb_scene_desc", "rtb_params", "rtb_ray", "rtb_hit", "rtb_shading", "rtb_stats"]
    src = '#include "rtb.h"\n#include <stdio.h>\nint main(){' + "".join(
        'printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names) + "return 0;}"
    c = tmp_path / "sz.c"
    c.write_text(src)
    exe = str(tmp_path / "sz")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", exe])
    sizes = dict(l.split() for l in subprocess.check_output([exe]).decode().splitlines())
    expect = {"rtb_camera": abi.camera_dt.itemsize, "rtb_ref_node": abi.ref_node_dt.itemsize,
              "rtb_tri_isect": abi.tri_isect_dt.itemsize, "rtb_tri_shade": abi.tri_shade_dt.itemsize,
              "rtb_material": abi.material_dt.itemsize, "rtb_texture": abi.texture_dt.itemsize,
              "rtb_light": abi.light_dt.itemsize, "rtb_scene_desc": C.sizeof(abi.SceneDesc),
              "rtb_params": C.sizeof(abi.Params), "rtb_ray": abi.ray_dt.itemsize, "rtb_hit": abi.hit_dt.itemsize,
              "rtb_shading": abi.shading_dt.itemsize, "rtb_stats": C.sizeof(abi.Stats)}
    assert {k: int(v) for k, v in sizes.items()} == expect


def test_default_params_are_the_reference_constants(rtb):
    from raytracingrenderer_b200 import abi
    p = rtb.default_params()
    assert p.max_depth == 4                       # Renderer.h:20
    assert p.epsilon == pytest.approx(1e-4)       # Geometry.h:60
    assert p.rr_cap == pytest.approx(0.9)         # Renderer.h:353
    assert p.integrator == abi.INT_PATH and p.filter == abi.FILTER_BOX and p.sampling == abi.SAMPLING_STRICT
    assert p.seed == 1 and p.partition == abi.PART_NONE


def test_no_cpu_fallback(rtb):
    """Without a CUDA device rtb_create must fail loudly (skipped where a GPU exists)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rtb.RtbError) as e:
        rtb.RayTracer(0)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product tree must not import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "raytracingrenderer_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|librtref|rtb_oracle", text):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_flat_scene_roundtrip(tmp_path):
    from raytracingrenderer_b200 import abi
    import refbvh
    s = refbvh.random_scene()
    p = str(tmp_path / "s.rtbs")
    s.save(p)
    t = abi.FlatScene.load(p)
    for k in ("ref_nodes", "tri_isect", "tri_shade", "materials", "textures", "lights", "texels"):
        assert getattr(s, k).tobytes() == getattr(t, k).tobytes(), k
    assert s.camera.tobytes() == t.camera.tobytes()
    assert (s.background_type, s.background_tex) == (t.background_type, t.background_tex)


def test_hdr_roundtrip(tmp_path):
    from raytracingrenderer_b200 import imageio
    rng = np.random.default_rng(0)
    img = (rng.random((17, 23, 3)) ** 4 * 50).astype(np.float32)
    p = str(tmp_path / "a.hdr")
    imageio.write_hdr(p, img)
    back = imageio.read_hdr(p)
    assert back.shape == img.shape
    # RGBE: 8-bit mantissa shared exponent
    assert np.all(np.abs(back - img) <= img.max(axis=-1, keepdims=True) / 128 + 1e-6)


def test_standalone_cpp_program_is_built_and_fails_loudly_without_a_gpu(rtb):
    """tools/rtb_render.cpp (Main.cpp's shape from this repository only) is part of the build; without a CUDA
    device it reports the library's error and exits non-zero."""
    import subprocess
    import torch
    from raytracingrenderer_b200 import build
    assert os.path.isfile(build.CLI), "run python -m raytracingrenderer_b200.build"
    out = subprocess.run([build.CLI], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "usage: rtb_render" in out.stderr
    scene = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes", "_staged", "cornell-box")
    if torch.cuda.is_available() or not os.path.isdir(scene):
        return
    out = subprocess.run([build.CLI, scene, "1", "/tmp/_rtb_render_test.hdr"], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
