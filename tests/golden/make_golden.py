#!/usr/bin/env python3
"""Regenerates the committed fixtures in tests/golden/ from the reference itself
(oracle/_ref/librtref.so, built by oracle/build_ref.py from /root/reference).  Run in the
build container only:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref  # noqa: E402
from raytracingrenderer_b200 import abi, imageio  # noqa: E402
import raysets  # noqa: E402

SCENES = ["cornell-box", "MaterialsScene", "materialball", "coffee", "bathroom"]


def fingerprints():
    out = {}
    for name in SCENES:
        s = ref.RefScene(name)
        ids, t = s.primary_hits()
        hit = ids != abi.MISS_ID
        out[name] = dict(width=s.width, height=s.height, n_tris=s.n_tris, n_lights=s.n_lights,
                         n_materials=s.n_materials,
                         ids_sha256_16=hashlib.sha256(ids.tobytes()).hexdigest()[:16],
                         t_sha256_16=hashlib.sha256(t.tobytes()).hexdigest()[:16],
                         sum_ids=int(ids[hit].astype(np.uint64).sum()), misses=int((~hit).sum()),
                         distinct_ids=int(len(np.unique(ids[hit]))))
        print(name, out[name])
    json.dump(out, open(os.path.join(HERE, "fingerprints.json"), "w"), indent=1)


def cornell():
    s = ref.RefScene("cornell-box")
    flat = s.flatten(os.path.join(HERE, "cornell-box.rtbs"))
    ids, t, rays = s.primary_hits(True)
    hits = s.trace(rays)
    sets = raysets.mixed_set(rays, hits, seed=20261018, n_each=400)
    closest_hits = s.trace(sets["closest"])
    any_hits = s.trace(sets["anyhit"], any_hit=True)
    vis = s.visible(sets["segments"])
    sd = s.shading_data(sets["closest"], closest_hits)
    ok = closest_hits["id"] != abi.MISS_ID
    rng = np.random.default_rng(7)
    n = int(ok.sum())
    wi = raysets.unit(rng.normal(size=(n, 3)))
    u = rng.random((n, 3), dtype=np.float32)
    b = s.eval_bsdf(sd[ok], wi, u)
    li = rng.integers(0, s.n_lights, n).astype(np.int32)
    L = s.eval_light(li, wi, u[:, :2])
    np.savez_compressed(os.path.join(HERE, "cornell_vectors.npz"), closest=sets["closest"], closest_hits=closest_hits,
                        anyhit=sets["anyhit"], anyhit_occluded=any_hits["id"].astype(np.uint8),
                        segments=sets["segments"], visible=vis, shading=sd, hit_mask=ok, wi=wi, u=u,
                        bsdf_eval=b["eval"], bsdf_pdf=b["pdf"], bsdf_s_wi=b["s_wi"], bsdf_s_f=b["s_f"],
                        bsdf_s_pdf=b["s_pdf"], light=li, light_p=L["p_or_wi"], light_emitted=L["emitted"],
                        light_pdf=L["pdf"], light_eval=L["eval"],
                        aov_albedo_blocks=raysets.block_mean(s.aov("albedo")),
                        aov_normals_blocks=raysets.block_mean(s.aov("normals")))
    # image statistics: two independent 32-spp halves of the reference renderer (all threads;
    # the reference is not deterministic across runs) + the committed 144-spp render
    a, na, _ = s.render(32, 0, fresh=True)
    bb, nb, _ = s.render(32, 0, fresh=False)  # continues the same RayTracer: film holds 64 spp
    bsum = bb - a
    r144 = imageio.read_hdr(os.path.join(ref.REF_DIR, "..", "..", "..", "reference", "RTBase", "result_144.hdr")) \
        if False else imageio.read_hdr("/root/reference/RTBase/result_144.hdr")
    np.savez_compressed(os.path.join(HERE, "cornell_ref_blocks.npz"),
                        half_a=raysets.block_mean(a / 32.0), half_b=raysets.block_mean(bsum / 32.0),
                        mean_a=(a / 32.0).mean(axis=(0, 1)), mean_b=(bsum / 32.0).mean(axis=(0, 1)),
                        result_144=raysets.block_mean(r144), result_144_mean=r144.mean(axis=(0, 1)),
                        pixel_rmse_halves=np.sqrt(np.mean((a / 32.0 - bsum / 32.0) ** 2)))
    print("cornell halves mean", (a / 32).mean(axis=(0, 1)), (bsum / 32).mean(axis=(0, 1)), "r144", r144.mean(axis=(0, 1)))
    cornell_mis()
    cornell_adaptive()
    cornell_light()
    cornell_ir()
    images()


def cornell_light():
    """RayTracer::lightTracer (Renderer.h:220-326) of the unmodified reference on cornell-box 256x256: two
    independent halves of 48 passes, 16x16-pixel block means."""
    s = ref.RefScene("cornell-box_256")
    a, _ = s.render_light(48, fresh=True)
    bb, _ = s.render_light(48, fresh=False)
    b = bb - a
    np.savez_compressed(os.path.join(HERE, "cornell256_light_blocks.npz"),
                        half_a=raysets.block_mean(a / 48.0, 16).astype(np.float32), half_b=raysets.block_mean(b / 48.0, 16).astype(np.float32),
                        mean_a=(a / 48.0).mean(axis=(0, 1)), mean_b=(b / 48.0).mean(axis=(0, 1)))
    print("cornell 256 light tracing halves mean", (a / 48).mean(axis=(0, 1)), (b / 48).mean(axis=(0, 1)))


def cornell_ir():
    """RayTracer::instantRadiosity (Renderer.h:82-218) of the unmodified reference on cornell-box 256x256: two
    independent halves of 64 passes (50 light paths each), 16x16-pixel block means."""
    s = ref.RefScene("cornell-box_256")
    a, _, na = s.render_ir(64, fresh=True)
    bb, _, nb = s.render_ir(64, fresh=False)
    b = bb - a
    np.savez_compressed(os.path.join(HERE, "cornell256_ir_blocks.npz"),
                        half_a=raysets.block_mean(a / 64.0, 16).astype(np.float32), half_b=raysets.block_mean(b / 64.0, 16).astype(np.float32),
                        mean_a=(a / 64.0).mean(axis=(0, 1)), mean_b=(b / 64.0).mean(axis=(0, 1)), vpls_per_pass=(na + nb) / 128.0)
    print("cornell 256 instant radiosity halves mean", (a / 64).mean(axis=(0, 1)), (b / 64).mean(axis=(0, 1)), "VPLs/pass", (na + nb) / 128.0)


def images():
    """Synthetic PNG / JPEG files (written with Pillow: baseline and progressive, 4:4:4 / 4:2:2 / 4:2:0, odd
    sizes, 1x1, optimised tables, restart intervals, grey) and the bytes the reference's stb_image decodes
    from them (oracle/_ref: ref_decode_image = stbi_load): the golden for the product's own decoders."""
    import hashlib, json
    from PIL import Image
    out = os.path.join(HERE, "images")
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(11)

    def pattern(w, h):
        y, x = np.mgrid[0:h, 0:w]
        img = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x * 7 + y * 13) % 256)], -1).astype(np.uint8)
        img[h // 4:h // 2, w // 4:w // 2] = rng.integers(0, 256, (h // 2 - h // 4, w // 2 - w // 4, 3), dtype=np.uint8)
        img[h // 2:, :w // 3] = [250, 10, 10]
        return img
    cases = []

    def jpg(name, w, h, **kw):
        Image.fromarray(pattern(w, h)).save(os.path.join(out, name), "JPEG", **kw)
        cases.append(name)
    jpg("base_444_q90.jpg", 64, 48, quality=90, subsampling=0)
    jpg("base_422_q75.jpg", 70, 50, quality=75, subsampling=1)
    jpg("base_420_q60_odd.jpg", 67, 45, quality=60, subsampling=2)
    jpg("base_420_q95_1px.jpg", 1, 1, quality=95, subsampling=2)
    jpg("base_420_17x9.jpg", 17, 9, quality=85, subsampling=2)
    jpg("prog_420_q80.jpg", 96, 80, quality=80, subsampling=2, progressive=True)
    jpg("prog_444_q50_odd.jpg", 53, 61, quality=50, subsampling=0, progressive=True)
    jpg("prog_422_q92.jpg", 40, 72, quality=92, subsampling=1, progressive=True)
    jpg("base_420_opt_q30.jpg", 128, 40, quality=30, subsampling=2, optimize=True)
    jpg("base_420_restart.jpg", 90, 40, quality=80, subsampling=2, restart_marker_blocks=3)
    jpg("prog_420_restart.jpg", 90, 40, quality=80, subsampling=2, progressive=True, restart_marker_rows=1)
    Image.fromarray(pattern(48, 40)[:, :, 0]).save(os.path.join(out, "grey_q80.jpg"), "JPEG", quality=80)
    cases.append("grey_q80.jpg")
    Image.fromarray(pattern(33, 29)).save(os.path.join(out, "rgb.png"), "PNG")
    cases.append("rgb.png")
    rgba = np.dstack([pattern(31, 17), (np.arange(31 * 17).reshape(17, 31) % 256).astype(np.uint8)])
    Image.fromarray(rgba, "RGBA").save(os.path.join(out, "rgba.png"), "PNG")
    cases.append("rgba.png")
    gold = {}
    for c in cases:
        a = ref.decode_image(os.path.join(out, c))
        gold[c] = {"shape": list(a.shape), "sha256": hashlib.sha256(a.tobytes()).hexdigest()}
    json.dump(gold, open(os.path.join(out, "stb_image_golden.json"), "w"), indent=1)
    print("images:", len(cases))


def cornell_adaptive():
    """RayTracer::adaptiveRender (Renderer.h:679-749) of the unmodified reference on cornell-box at
    256x256 (scenes/_staged/cornell-box_256: the bundled scene.json with width/height edited)."""
    s = ref.RefScene("cornell-box_256")
    s.flatten(os.path.join(HERE, "cornell-box_256.rtbs"))
    film, cnt, var, secs = s.render_adaptive()
    np.savez_compressed(os.path.join(HERE, "cornell256_adaptive.npz"), tile_samples=cnt, tile_variance=var,
                        film_blocks=raysets.block_mean(film, 8).astype(np.float32), film_mean=film.mean(axis=(0, 1)))
    print("cornell 256 adaptive: %.1f s, mean count %.1f, film mean" % (secs, cnt.mean()), film.mean(axis=(0, 1)))


def cornell_mis():
    """pathTrace with computeDirectMIS (Renderer.h:474-557; oracle/_ref/librtref_mis.so): two
    independent 16-spp halves, 16x16-pixel block means."""
    m = ref.RefScene("cornell-box", "_mis")
    a, _, _ = m.render(16, 0, fresh=True)
    bb, _, _ = m.render(16, 0, fresh=False)
    b = bb - a
    np.savez_compressed(os.path.join(HERE, "cornell_mis_blocks.npz"),
                        half_a=raysets.block_mean(a / 16.0, 16).astype(np.float32),
                        half_b=raysets.block_mean(b / 16.0, 16).astype(np.float32),
                        mean_a=(a / 16.0).mean(axis=(0, 1)), mean_b=(b / 16.0).mean(axis=(0, 1)))
    print("cornell MIS halves mean", (a / 16).mean(axis=(0, 1)), (b / 16).mean(axis=(0, 1)))


def all_bsdfs():
    """BSDF vectors over every material class the loader can produce (MaterialsScene has
    mirror/diffuse/plastic/conductor/orennayar/glass; the materialball override adds the
    rough dielectric), with the material + texture tables needed to re-create them."""
    mats, texs, texels, sds, wis, us, outs = [], [], [], [], [], [], []
    rng = np.random.default_rng(11)
    for name in ["MaterialsScene", "materialball_dielectric", "materialball_glass"]:
        s = ref.RefScene(name)
        flat = s.flatten("/tmp/_golden_%s.rtbs" % name)
        ids, t, rays = s.primary_hits(True)
        hits = s.trace(rays)
        pts, m = raysets.hit_points(rays, hits)
        # primary + one random bounce (inside-glass paths need rays that start inside)
        sub = rng.choice(np.where(m)[0], 1500, replace=False)
        br = raysets.bounce_rays(pts, rng, 1500)
        allr = np.concatenate([rays[sub], br])
        allh = s.trace(allr)
        ok = allh["id"] != abi.MISS_ID
        sd = s.shading_data(allr[ok], allh[ok])
        # keep <= 120 records per material
        keep = []
        for mi in np.unique(sd["material"]):
            w = np.where(sd["material"] == mi)[0]
            keep.extend(w[:120])
        sd = sd[np.array(sorted(keep))]
        n = len(sd)
        wi = raysets.unit(rng.normal(size=(n, 3)))
        u = rng.random((n, 3), dtype=np.float32)
        o = s.eval_bsdf(sd, wi, u)
        # re-base material/texture indices into the concatenated tables; only material
        # textures (1x1 PNGs) are kept, never the env map
        mat_tex = flat.materials["tex"]
        used_tex = sorted(set(int(x) for x in mat_tex))
        tex_map = {}
        for ti in used_tex:
            T = flat.textures[ti]
            px = flat.texels.reshape(-1, 3)[T["offset"]:T["offset"] + T["width"] * T["height"]]
            nt = np.zeros((), abi.texture_dt)
            nt["offset"], nt["width"], nt["height"] = sum(len(x) for x in texels), T["width"], T["height"]
            tex_map[ti] = len(texs)
            texs.append(nt)
            texels.append(px.copy())
        base = len(mats)
        for mrec in flat.materials:
            mrec = mrec.copy()
            mrec["tex"] = tex_map[int(mrec["tex"])]
            mats.append(mrec)
        sd = sd.copy()
        sd["material"] += base
        sds.append(sd), wis.append(wi), us.append(u), outs.append(o)
    np.savez_compressed(os.path.join(HERE, "bsdf_vectors.npz"), materials=np.array(mats, abi.material_dt),
                        textures=np.array(texs, abi.texture_dt), texels=np.concatenate(texels).astype(np.float32),
                        shading=np.concatenate(sds), wi=np.concatenate(wis), u=np.concatenate(us),
                        **{k: np.concatenate([o[k] for o in outs]) for k in outs[0]})
    types = np.array(mats, abi.material_dt)["type"]
    print("bsdf vectors:", sum(len(x) for x in sds), "records; material types", sorted(set(int(x) for x in types)))


if __name__ == "__main__":
    what = sys.argv[1:] or ["fingerprints", "cornell", "bsdfs"]
    if "fingerprints" in what:
        fingerprints()
    if "cornell" in what:
        cornell()
    if "bsdfs" in what:
        all_bsdfs()
