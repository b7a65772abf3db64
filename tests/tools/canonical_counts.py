"""SURVEY 8(d)'s work counter: replays every ray of a 1-spp render of each BASELINE scene through the
CANONICAL traversal of the reference tree (oracle/rtb_oracle.c: canonical_count) and writes the per-ray
means to profiles/canonical_counts.json.  bench.py computes roofline.achieved from these figures
(bytes/ray = 32 n_box + 64 n_tri + 48, flops/ray = 24 n_box + 60 n_tri), whatever tree the kernels walk.
CPU only; run here:  python tests/tools/canonical_counts.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from raytracingrenderer_b200 import host_api
from oracle import port

VARIANTS = ["diffuse", "conductor", "glass", "dielectric", "orennayar", "plastic", "layered"]
scenes = ["cornell-box", "materialball", "MaterialsScene", "MaterialsScene_env", "coffee", "bathroom"] + \
         ["materialball_" + v for v in VARIANTS]
out = {}
for name in scenes:
    s = host_api.load_scene(os.path.join(ROOT, "scenes", "_staged", name))
    _, st = port.Oracle(s).render_counts(1)
    c, sh, n = st["closest_rays"], st["shadow_rays"], st["samples"]
    row = {"closest_box": st["closest_box"] / c, "closest_tri": st["closest_tri"] / c,
           "shadow_box": st["shadow_box"] / max(sh, 1), "shadow_tri": st["shadow_tri"] / max(sh, 1),
           "closest_rays_per_sample": c / n, "shadow_rays_per_sample": sh / n, "spp": 1, "triangles": s.n_tris}
    row["closest_bytes_per_ray"] = 32 * row["closest_box"] + 64 * row["closest_tri"] + 48
    row["shadow_bytes_per_ray"] = 32 * row["shadow_box"] + 64 * row["shadow_tri"] + 48
    row["closest_flops_per_ray"] = 24 * row["closest_box"] + 60 * row["closest_tri"]
    row["shadow_flops_per_ray"] = 24 * row["shadow_box"] + 60 * row["shadow_tri"]
    out[name] = row
    print(name, json.dumps(row), flush=True)
for lg in (20,):
    s, _ = host_api.build_soup(1 << lg, 3840, 2160)
    _, st = port.Oracle(s, max_depth=0).render_counts(1)
    c, sh, n = st["closest_rays"], st["shadow_rays"], st["samples"]
    row = {"closest_box": st["closest_box"] / c, "closest_tri": st["closest_tri"] / c,
           "shadow_box": st["shadow_box"] / max(sh, 1), "shadow_tri": st["shadow_tri"] / max(sh, 1),
           "closest_rays_per_sample": c / n, "shadow_rays_per_sample": sh / n, "spp": 1, "triangles": s.n_tris}
    row["closest_bytes_per_ray"] = 32 * row["closest_box"] + 64 * row["closest_tri"] + 48
    row["shadow_bytes_per_ray"] = 32 * row["shadow_box"] + 64 * row["shadow_tri"] + 48
    out["soup%d" % lg] = row
    print("soup", lg, json.dumps(row), flush=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "canonical_counts.json"), "w"), indent=1)
