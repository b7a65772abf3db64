"""Parity tests proper: the CUDA path, called through the C ABI (librtb200.so), against the
oracle (C restatement, always) and the unmodified reference (oracle/_ref, when it travelled to
the box) on the same seeded inputs.  Bars: hit IDs / t / barycentrics bit-exact; shading data,
BSDF and light evaluations <= 1e-5 relative; images statistically."""
import hashlib
import json
import os

import numpy as np
import pytest

import raysets
from conftest import BUNDLED, GOLDEN, flat_scene, ref_scene, rel_err, synthetic_scene
from raytracingrenderer_b200 import abi

pytestmark = pytest.mark.gpu

FP = json.load(open(os.path.join(GOLDEN, "fingerprints.json")))
TRAVS = [abi.TRAV_EXACT, abi.TRAV_FAST, abi.TRAV_WIDE, abi.TRAV_CW, abi.TRAV_Q16]


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


_rt_cache = {}


def gpu_scene(rtb, name):
    """One RayTracer per scene for the whole session (upload once)."""
    if name not in _rt_cache:
        rt = rtb.RayTracer(0)
        rt.init(synthetic_scene() if name == "synthetic" else flat_scene(name))
        _rt_cache[name] = rt
    rt = _rt_cache[name]
    p = rtb.default_params()
    rt.set_params(**{k: getattr(p, k) for k, _ in abi.Params._fields_})
    rt.clear()
    return rt


# ------------------------------------------------------------------ boundary behaviour
def test_errors_are_reported_not_swallowed(rtb):
    rt = rtb.RayTracer(0)
    with pytest.raises(rtb.RtbError) as e:
        rt.render(1, 0)
    assert e.value.code == -3                      # RTB_ERR_STATE: render before upload
    with pytest.raises(rtb.RtbError) as e:
        rt.set_params(integrator=17)
    assert e.value.code == -1
    with pytest.raises(rtb.RtbError):
        rtb.RayTracer(4096)                        # no such device
    s = synthetic_scene()
    s.tri_isect = s.tri_isect.copy()
    s.tri_isect["material"][0] = 999
    with pytest.raises(rtb.RtbError) as e:
        rt.init(s)
    assert "material" in str(e.value)
    # a texture that runs past the texel pool, and a node array that is not one pre-order tree (the skip links of the
    # EXACT traversal would walk out of it): errors at upload, not reads out of bounds on the device
    s = synthetic_scene()
    s.textures = s.textures.copy()
    s.textures["width"][0] = 1 << 14
    s.textures["height"][0] = 1 << 14
    with pytest.raises(rtb.RtbError) as e:
        rt.init(s)
    assert "texel pool" in str(e.value)
    s = synthetic_scene()
    s.ref_nodes = s.ref_nodes.copy()
    interior = np.flatnonzero(s.ref_nodes["a"] >= 0)
    s.ref_nodes["b"][interior[len(interior) // 2]] = len(s.ref_nodes) + 5
    with pytest.raises(rtb.RtbError) as e:
        rt.init(s)
    assert "pre-order" in str(e.value)
    rt.init(synthetic_scene())                     # the context is still usable after the rejected uploads
    rt.render(1, 0)
    assert np.isfinite(rt.read_film()).all()
    rt.close()


def test_parameters_that_would_render_wrong_images_are_rejected(rtb):
    """A zero-initialised rtb_params (instead of rtb_default_params) has rr_cap = 0 (every path ends at depth 0) and
    cull_rel = 0 (no cull slack: FAST / WIDE hit IDs could differ); films above 2^32 / 3 pixels would be resolved only
    partially.  All of these are errors, not silently wrong images."""
    rt = rtb.RayTracer(0)
    for kw in (dict(rr_cap=0.0), dict(rr_cap=1.5), dict(rr_cap=float("nan")), dict(cull_rel=0.0), dict(cull_rel=-1e-5),
               dict(cull_rel=float("nan")), dict(filter_radius=float("nan")), dict(filter_alpha=float("inf")), dict(primary_reuse=7),
               dict(traversal=9)):
        with pytest.raises(rtb.RtbError) as e:
            rt.set_params(**kw)
        assert e.value.code == -1, kw
    rt.set_params(rr_cap=1.0, cull_rel=1e-6)              # the edges of the accepted ranges
    s = synthetic_scene()
    big = abi.FlatScene()
    big.__dict__.update(s.__dict__)
    big.camera = s.camera.copy()
    big.camera["width"], big.camera["height"] = 40000.0, 40000.0
    with pytest.raises(rtb.RtbError) as e:
        rt.init(big)
    assert "too large" in str(e.value)
    rt.close()


def test_rng_stream_equals_the_oracle(rtb, oracle_mod):
    rt = gpu_scene(rtb, "synthetic")
    for seed, pixel, sample in ((1, 0, 0), (1, 12345, 77), (0xB200, 2 ** 31 + 5, 2 ** 20)):
        rt.set_params(seed=seed)
        assert np.array_equal(rt.rng_draws(pixel, sample, 48), oracle_mod.rng_draws(seed, pixel, sample, 48))


# ------------------------------------------------------------------ gate 1: hit IDs bit-exact
@pytest.mark.parametrize("trav", TRAVS)
def test_cornell_primary_hits_golden(rtb, trav):
    rt = gpu_scene(rtb, "cornell-box")
    ids, t = rt.primary_hits(trav)
    assert sha16(ids) == FP["cornell-box"]["ids_sha256_16"] == "0e060cc6996f198b"
    assert sha16(t) == FP["cornell-box"]["t_sha256_16"] == "1227c2cf5a6229a9"
    assert int((ids == abi.MISS_ID).sum()) == 19


@pytest.mark.parametrize("name", BUNDLED[1:])
@pytest.mark.parametrize("trav", TRAVS)
def test_primary_hits_bit_exact_all_scenes(rtb, name, trav):
    rt = gpu_scene(rtb, name)          # skips when oracle/_ref did not travel
    ids, t, rays = rt.primary_hits(trav, want_rays=True)
    assert sha16(ids) == FP[name]["ids_sha256_16"]
    assert sha16(t) == FP[name]["t_sha256_16"]
    rid, rtt, rrays = ref_scene(name).primary_hits(want_rays=True)
    assert np.array_equal(ids, rid) and t.tobytes() == rtt.tobytes()
    assert rays["o"].tobytes() == rrays["o"].tobytes() and rays["d"].tobytes() == rrays["d"].tobytes()


@pytest.mark.parametrize("trav", TRAVS)
def test_cornell_golden_ray_sets(rtb, trav):
    """Recorded rays incl. the axis-aligned 0*inf = NaN case."""
    rt = gpu_scene(rtb, "cornell-box")
    cv = np.load(os.path.join(GOLDEN, "cornell_vectors.npz"))
    assert rt.trace(cv["closest"], traversal=trav).tobytes() == cv["closest_hits"].tobytes()
    occ = rt.trace(cv["anyhit"], any_hit=True, traversal=trav)["id"].astype(np.uint8)
    assert np.array_equal(occ, cv["anyhit_occluded"])
    assert np.array_equal(rt.visible(cv["segments"], traversal=trav), cv["visible"])


@pytest.mark.parametrize("name", ["synthetic", "cornell-box", "MaterialsScene", "materialball", "coffee", "bathroom"])
def test_all_ray_kinds_equal_the_oracle(rtb, oracle_mod, name):
    """Closest-hit, any-hit and Scene::visible on primary + bounce + axis-aligned rays:
    EXACT == FAST == oracle, bit for bit."""
    rt = gpu_scene(rtb, name)
    o = oracle_mod.Oracle(rt.scene)
    ids, t, rays = rt.primary_hits(abi.TRAV_EXACT, want_rays=True)
    ph = rt.trace(rays, traversal=abi.TRAV_EXACT)
    n_each = 20000 if name in ("synthetic", "cornell-box") else 3000
    sets = raysets.mixed_set(rays, ph, seed=4, n_each=n_each)
    want = o.trace(sets["closest"])
    want_any = o.trace(sets["anyhit"], any_hit=True)["id"]
    want_vis = o.visible(sets["segments"])
    for trav in TRAVS:
        assert rt.trace(sets["closest"], traversal=trav).tobytes() == want.tobytes(), trav
        assert np.array_equal(rt.trace(sets["anyhit"], any_hit=True, traversal=trav)["id"], want_any), trav
        assert np.array_equal(rt.visible(sets["segments"], traversal=trav), want_vis), trav


@pytest.mark.parametrize("name", ["cornell-box", "coffee"])
def test_nan_rays_behave_like_the_reference_and_cost_nothing(rtb, oracle_mod, name):
    """A NaN in a ray's origin or direction (the reference's glass / normalise code can produce one) can
    never be accepted as a closest hit (t is NaN for every triangle) and is "occluded" by the first triangle
    an any-hit walk tests.  Same answers as the oracle's exhaustive walk — without walking the whole tree
    (one such ray in bathroom sample 861 used to hold a persistent warp for 0.35 s)."""
    import time
    rt = gpu_scene(rtb, name)
    o = oracle_mod.Oracle(rt.scene)
    _, _, rays = rt.primary_hits(abi.TRAV_EXACT, want_rays=True)
    rng = np.random.default_rng(9)
    base = rays[rng.choice(len(rays), 600, replace=False)].copy()
    for k, comp in enumerate((("o", 0), ("o", 1), ("o", 2), ("d", 0), ("d", 1), ("d", 2))):
        base[comp[0]][k * 100:(k + 1) * 100, comp[1]] = np.nan
    want = o.trace(base)
    assert np.all(want["id"] == abi.MISS_ID)
    want_any = o.trace(base, any_hit=True)["id"]
    big = np.tile(base, (400, 1))                   # 240 k NaN rays: seconds if each walked the tree
    for trav in TRAVS:
        got = rt.trace(base, traversal=trav)
        assert np.array_equal(got["id"], want["id"]) and got["t"].tobytes() == want["t"].tobytes(), trav
        assert np.array_equal(rt.trace(base, any_hit=True, traversal=trav)["id"], want_any), trav
        t0 = time.time()
        rt.trace(big, traversal=trav)
        assert time.time() - t0 < 1.0, trav


# ------------------------------------------------------------------ gate 2: evaluations <= 1e-5
def test_cornell_golden_shading_bsdf_light(rtb):
    rt = gpu_scene(rtb, "cornell-box")
    cv = np.load(os.path.join(GOLDEN, "cornell_vectors.npz"))
    m = cv["hit_mask"]
    sd = rt.shading_data(cv["closest"], cv["closest_hits"])
    for f in ("x", "wo", "s_normal", "g_normal", "frame_u", "frame_v", "frame_w"):
        assert rel_err(sd[f][m], cv["shading"][f][m]) <= 1e-5, f
    for f in ("tu", "tv", "t"):
        assert rel_err(sd[f][m], cv["shading"][f][m]) <= 1e-5, f
    assert np.array_equal(sd["material"][m], cv["shading"]["material"][m])
    assert np.all(sd["material"][~m] == -1) and sd["wo"][~m].tobytes() == cv["shading"]["wo"][~m].tobytes()
    b = rt.eval_bsdf(cv["shading"][m], cv["wi"], cv["u"])
    for k in ("eval", "pdf", "s_wi", "s_f", "s_pdf"):
        assert rel_err(b[k], cv["bsdf_" + k]) <= 1e-5, k
    L = rt.eval_light(cv["light"], cv["wi"], cv["u"][:, :2])
    for k, g in (("p_or_wi", "light_p"), ("emitted", "light_emitted"), ("pdf", "light_pdf"), ("eval", "light_eval")):
        assert rel_err(L[k], cv[g]) <= 1e-5, k


def test_glass_fresnel_nan_reflects_like_the_reference(rtb, oracle_mod):
    """|cos_i| a hair above 1 (wo antiparallel to a rounded normal) makes sqrtf(1 - cos^2) NaN inside
    ShadingHelper::fresnelDielectric; the reference's clamp = std::max(0, std::min(1, NaN)) is 1, i.e. total
    reflection.  (nvcc fuses that select pattern into FMUL.SAT, which flushes NaN to 0: the GPU used to refract
    along a NaN direction — the NaN rays of profiles/r01_scaling.md.)"""
    rt = gpu_scene(rtb, "synthetic")
    s = rt.scene
    glass = [i for i, m in enumerate(s.materials) if m["type"] == abi.BSDF_GLASS]
    assert glass
    N = 256
    rng = np.random.default_rng(3)
    n = rng.normal(size=(N, 3))
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    a = np.where(np.abs(n[:, :1]) > 0.9, [[0, 1, 0]], [[1, 0, 0]]).astype(np.float32)
    u = np.cross(n, a)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = np.cross(n, u)
    sd = np.zeros(N, abi.shading_dt)
    sd["s_normal"], sd["g_normal"], sd["frame_u"], sd["frame_v"], sd["frame_w"] = n, n, u, v, n
    sd["t"], sd["material"] = 1, glass[0]
    wi = np.tile(np.array([[0, 1, 0]], np.float32), (N, 1))
    uu = rng.random((N, 3), dtype=np.float32)
    o = oracle_mod.Oracle(s)
    for scale in (1.0000002, 1.0000005):
        for sign in (1.0, -1.0):
            sd["wo"] = (n * np.float32(sign * scale)).astype(np.float32)
            want = o.eval_bsdf(sd, wi, uu)
            assert np.all(want["s_pdf"] == 1.0) and np.isfinite(want["s_wi"]).all()       # the reference reflects
            got = rt.eval_bsdf(sd, wi, uu)
            assert np.isfinite(got["s_wi"]).all()
            for k in ("s_wi", "s_f", "s_pdf"):
                assert rel_err(got[k], want[k]) <= 1e-5, (k, scale, sign)


def test_every_bsdf_class_golden(rtb):
    g = np.load(os.path.join(GOLDEN, "bsdf_vectors.npz"))
    s = flat_scene("cornell-box")
    t = abi.FlatScene()
    t.camera, t.ref_nodes, t.tri_isect, t.tri_shade = s.camera, s.ref_nodes, s.tri_isect.copy(), s.tri_shade
    t.tri_isect["material"] = 0
    t.materials, t.textures, t.texels = g["materials"], g["textures"], g["texels"].ravel()
    rt = rtb.RayTracer(0)
    rt.init(t)
    b = rt.eval_bsdf(g["shading"], g["wi"], g["u"])
    types = g["materials"]["type"][g["shading"]["material"]]
    for ty in range(7):
        sel = types == ty
        assert sel.sum() > 20
        for k in ("eval", "pdf", "s_wi", "s_f", "s_pdf"):
            assert rel_err(b[k][sel], g[k][sel]) <= 1e-5, (ty, k)
    rt.close()


@pytest.mark.parametrize("name", ["synthetic", "MaterialsScene", "materialball", "bathroom", "materialball_dielectric"])
def test_shading_bsdf_light_equal_the_oracle(rtb, oracle_mod, name):
    rt = gpu_scene(rtb, name)
    o = oracle_mod.Oracle(rt.scene)
    ids, t, rays = rt.primary_hits(abi.TRAV_EXACT, want_rays=True)
    ph = rt.trace(rays, traversal=abi.TRAV_EXACT)
    rng = np.random.default_rng(8)
    pts, _ = raysets.hit_points(rays, ph)
    rr = np.concatenate([rays[rng.integers(0, len(rays), 4000)], raysets.bounce_rays(pts, rng, 4000)])
    hits = o.trace(rr)
    m = hits["id"] != abi.MISS_ID
    sd_o, sd_g = o.shading_data(rr, hits), rt.shading_data(rr, hits)
    for f in ("x", "wo", "s_normal", "g_normal", "frame_u", "frame_v", "frame_w", "tu", "tv", "t"):
        assert rel_err(sd_g[f][m], sd_o[f][m]) <= 1e-5, f
    assert np.array_equal(sd_g["material"], sd_o["material"])
    n = int(m.sum())
    wi, u = raysets.unit(rng.normal(size=(n, 3))), rng.random((n, 3), dtype=np.float32)
    a, b = o.eval_bsdf(sd_o[m], wi, u), rt.eval_bsdf(sd_o[m], wi, u)
    for k in a:
        assert rel_err(b[k], a[k]) <= 1e-5, k
    nl = len(rt.scene.lights)
    li = rng.integers(0, nl, n).astype(np.int32)
    a, b = o.eval_light(li, wi, u[:, :2]), rt.eval_light(li, wi, u[:, :2])
    assert rel_err(b["p_or_wi"], a["p_or_wi"]) <= 1e-5 and rel_err(b["pdf"], a["pdf"]) <= 1e-5
    env = rt.scene.lights["type"][li] == abi.LIGHT_ENVMAP
    for k in ("emitted", "eval"):
        assert rel_err(b[k][~env], a[k][~env]) <= 1e-5, k
    if env.any():
        # An env-map lookup magnifies a few-ulp difference of atan2f/acosf (device libm vs
        # glibc) by width x the local texel gradient.  Allowance: 1e-5 relative plus a texel
        # coordinate uncertainty of 1e-3 texel times the local 3x3 texel range.
        tex = int(rt.scene.lights["tex"][li[env][0]])
        want = o.eval_light(li[env], wi[env], u[env][:, :2])["eval"]
        tol = 1e-5 * np.abs(want) + raysets.env_lookup_slack(rt.scene, tex, wi[env], 1e-3) + 1e-12
        assert np.all(np.abs(b["eval"][env] - want) <= tol)
        # emitted = lookup at the sampled direction: judged at the GPU's own direction
        gdir = np.ascontiguousarray(b["p_or_wi"][env])
        want = o.eval_light(li[env], gdir, u[env][:, :2])["eval"]
        tol = 1e-5 * np.abs(want) + raysets.env_lookup_slack(rt.scene, tex, gdir, 1e-3) + 1e-12
        assert np.all(np.abs(b["emitted"][env] - want) <= tol)


@pytest.mark.parametrize("name", ["cornell-box", "MaterialsScene", "bathroom"])
def test_aov_integrators(rtb, oracle_mod, name):
    """albedo / viewNormals (Renderer.h:558-581): deterministic, so <= 1e-5 everywhere."""
    rt = gpu_scene(rtb, name)
    for integ, kind in ((abi.INT_ALBEDO, "albedo"), (abi.INT_NORMALS, "normals")):
        rt.set_params(integrator=integ)
        rt.clear()
        rt.render(1, 0)
        img = rt.read_film()
        want, _ = oracle_mod.Oracle(rt.scene, integrator=integ).render(1)
        assert np.allclose(img, want, rtol=1e-5, atol=1e-6), kind
        if name != "cornell-box" or True:
            try:
                assert np.allclose(img, ref_scene(name).aov(kind), rtol=1e-5, atol=1e-6), kind
            except pytest.skip.Exception:
                pass
    cv = np.load(os.path.join(GOLDEN, "cornell_vectors.npz"))
    if name == "cornell-box":
        rt.set_params(integrator=abi.INT_NORMALS)
        rt.clear()
        rt.render(1, 0)
        assert np.allclose(raysets.block_mean(rt.read_film()), cv["aov_normals_blocks"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["cornell-box", "coffee", "materialball"])
def test_direct_integrator_equals_the_oracle_and_the_reference(rtb, oracle_mod, name):
    """RTB_INT_DIRECT = RayTracer::direct (Renderer.h:393-407): closest hit, emission if the hit is a light,
    else ONE computeDirect sample; a miss is black.  Against the oracle the uniforms are the same (block 0 of
    the pixel sample): sample for sample.  Against the reference's own direct() (1 sample per pixel with its
    MTRandom): 8x8-block means within the estimator's noise, which is measured from two GPU runs with
    disjoint sample indices — RMSE(GPU N spp vs reference 1 spp) must sit at sigma_1 sqrt(1 + 1/N), +-20 %."""
    rt = gpu_scene(rtb, name)
    rt.set_params(integrator=abi.INT_DIRECT, primary_reuse=0)
    rt.render(2, 0)
    img = rt.read_film().copy()
    g = rt.stats()
    want, st = oracle_mod.Oracle(rt.scene, integrator=abi.INT_DIRECT).render(2)
    close = np.isclose(img, want, rtol=2e-4, atol=1e-5).all(axis=-1)
    assert close.mean() > 0.995, close.mean()
    assert g["samples"] == st["samples"] == rt.width * rt.height * 2
    assert g["closest_rays"] == st["closest_rays"] == g["samples"]           # one camera ray per sample, nothing else
    assert abs(g["shadow_rays"] / max(st["shadow_rays"], 1) - 1) < 2e-3
    # all three traversals and the primary-hit table give the same bits
    for kw in (dict(traversal=abi.TRAV_EXACT), dict(traversal=abi.TRAV_WIDE), dict(traversal=abi.TRAV_CW), dict(traversal=abi.TRAV_Q16), dict(primary_reuse=1)):
        rt.set_params(integrator=abi.INT_DIRECT, traversal=abi.TRAV_FAST, primary_reuse=0)
        rt.set_params(**kw)
        rt.clear()
        rt.render(2, 0)
        assert rt.read_film().tobytes() == img.tobytes(), kw
    # the megakernel schedule: same samples
    rt.set_params(traversal=abi.TRAV_FAST, primary_reuse=0, scheduler=abi.SCHED_MEGAKERNEL)
    rt.clear()
    rt.render(2, 0)
    assert np.allclose(rt.read_film(), img, rtol=1e-5, atol=1e-6)
    rt.set_params(scheduler=abi.SCHED_WAVEFRONT)
    try:
        ref1 = ref_scene(name).aov("direct")
    except pytest.skip.Exception:
        return
    N = 16
    rt.clear()
    rt.render(1, 100)
    a = rt.read_film().copy()
    rt.clear()
    rt.render(1, 200)
    b = rt.read_film().copy()
    rt.clear()
    rt.render(N, 0)
    mean = rt.read_film() / N
    ba, bb, bm, br = (raysets.block_mean(x) for x in (a, b, mean, ref1))
    sigma1 = np.sqrt(np.mean((ba - bb) ** 2) / 2)           # block noise of one sample per pixel
    rmse = np.sqrt(np.mean((bm - br) ** 2))
    expect = sigma1 * np.sqrt(1 + 1 / N)
    assert 0.8 * expect < rmse < 1.2 * expect, (rmse, expect)
    assert np.all(np.abs(mean.mean(axis=(0, 1)) / ref1.mean(axis=(0, 1)) - 1) < 0.02)


# ------------------------------------------------------------------ gate 3: images
@pytest.mark.parametrize("name", ["synthetic", "cornell-box", "materialball"])
def test_render_equals_the_oracle_sample_for_sample(rtb, oracle_mod, name):
    """Same counter-based RNG on both sides: every pixel sum must agree except where a libm
    ulp (acosf / sincosf / atan2f) flips a discrete path decision."""
    rt = gpu_scene(rtb, name)
    spp = 2
    rt.set_params(primary_reuse=0)        # every sample traces its own camera ray, like the reference
    rt.render(spp, 0)
    img = rt.read_film()
    want, st = oracle_mod.Oracle(rt.scene).render(spp)
    assert not np.isnan(img).any()
    close = np.isclose(img, want, rtol=2e-4, atol=1e-5).all(axis=-1)
    assert close.mean() > 0.995, close.mean()
    assert np.all(np.abs(img.mean(axis=(0, 1)) / want.mean(axis=(0, 1)) - 1) < 5e-3)
    g = rt.stats()
    assert g["samples"] == st["samples"] == rt.width * rt.height * spp
    assert abs(g["closest_rays"] / st["closest_rays"] - 1) < 2e-3
    assert abs(g["shadow_rays"] / st["shadow_rays"] - 1) < 2e-3


@pytest.mark.parametrize("name", ["synthetic", "cornell-box", "materialball"])
def test_mis_estimator_equals_the_oracle_sample_for_sample(rtb, oracle_mod, name):
    """RTB_INT_PATH_MIS (computeDirectMIS, Renderer.h:474-557): same uniforms on both sides."""
    rt = gpu_scene(rtb, name)
    spp = 2
    rt.set_params(integrator=abi.INT_PATH_MIS, primary_reuse=0)
    rt.render(spp, 0)
    img = rt.read_film()
    want, st = oracle_mod.Oracle(rt.scene, integrator=abi.INT_PATH_MIS).render(spp)
    assert not np.isnan(img).any()
    close = np.isclose(img, want, rtol=2e-4, atol=1e-5).all(axis=-1)
    assert close.mean() > 0.995, close.mean()
    assert np.all(np.abs(img.mean(axis=(0, 1)) / want.mean(axis=(0, 1)) - 1) < 5e-3)
    g = rt.stats()
    assert g["samples"] == st["samples"]
    assert abs(g["closest_rays"] / st["closest_rays"] - 1) < 2e-3
    assert abs(g["shadow_rays"] / st["shadow_rays"] - 1) < 2e-3
    # the megakernel schedule traces the probe ray in place: same samples, same counters
    rt.set_params(scheduler=abi.SCHED_MEGAKERNEL)
    rt.clear()
    rt.render(spp, 0)
    mk = rt.read_film()
    assert np.allclose(mk, img, rtol=1e-4, atol=1e-5)
    gm_ = rt.stats()
    assert (gm_["closest_rays"], gm_["shadow_rays"]) == (g["closest_rays"], g["shadow_rays"])
    rt.set_params(scheduler=abi.SCHED_WAVEFRONT, primary_reuse=1)
    rt.clear()
    rt.render(spp, 0)
    assert np.array_equal(rt.read_film(), img)           # primary-hit table: same bits
    if name == "cornell-box":
        gm = np.load(os.path.join(GOLDEN, "cornell_mis_blocks.npz"))
        rt.clear()
        rt.render(16, 0)
        m = (rt.read_film() / 16).mean(axis=(0, 1))
        assert np.all(np.abs(m / (0.5 * (gm["mean_a"] + gm["mean_b"])) - 1) < 0.01)


def test_adaptive_render_equals_the_oracle_and_the_reference(rtb, oracle_mod):
    """rtb_render_adaptive = RayTracer::adaptiveRender (Renderer.h:583-749).  Against the oracle the
    uniforms are identical: tile variances to 1e-3, sample counts equal up to one sample where the
    float rounding of the variance sums differs, pixels of equally-sampled tiles sample for sample.
    Against the reference's own run (golden): the statistics of test_oracle_cpu."""
    g = np.load(os.path.join(GOLDEN, "cornell256_adaptive.npz"))
    s = abi.FlatScene.load(os.path.join(GOLDEN, "cornell-box_256.rtbs"))
    rt = rtb.RayTracer(0)
    rt.init(s)
    cnt, var = rt.adaptiveRender(2, 1, 10240)
    img = rt.read_film()
    assert rt.getSPP() == 1
    want, ocnt, ovar = oracle_mod.Oracle(s).render_adaptive(2, 1, 10240)
    assert np.allclose(var, ovar, rtol=1e-3, atol=1e-7)
    d = np.abs(cnt.astype(np.int64) - ocnt.astype(np.int64))
    assert d.max() <= 1 and (d == 0).mean() > 0.9, (d.max(), (d == 0).mean())
    same = np.kron(d == 0, np.ones((32, 32), bool))[:rt.height, :rt.width]
    close = np.isclose(img, want, rtol=5e-4, atol=1e-5).all(axis=-1)
    assert close[same].mean() > 0.99, close[same].mean()
    assert np.all(np.abs(img.mean(axis=(0, 1)) / want.mean(axis=(0, 1)) - 1) < 5e-3)
    st = rt.stats()
    assert st["samples"] == 256 * 256 * 2 + int(cnt.astype(np.int64).sum()) * 1024
    # the reference itself
    big = g["tile_variance"] > 0.01 * g["tile_variance"].max()
    assert np.all(np.abs(var[big] / g["tile_variance"][big] - 1) < 0.02)
    assert np.all(np.abs(cnt[big].astype(float) / g["tile_samples"][big] - 1) < 0.01)
    assert np.all(np.abs(img.mean(axis=(0, 1)) / g["film_mean"] - 1) < 0.01)
    # a second call adds a second, INDEPENDENT mean image (the reference's MTRandom keeps advancing between
    # render() calls): it draws the sample indices [(2 + 10240), 2 (2 + 10240)), checked against the oracle
    cnt2, _ = rt.adaptiveRender(2, 1, 10240)
    assert rt.getSPP() == 2
    both = rt.read_film().copy()
    assert np.all(np.abs(both.mean(axis=(0, 1)) / (2 * g["film_mean"]) - 1) < 0.01)
    second = both - img
    assert (np.abs(second - img) > 1e-6).any(axis=-1).mean() > 0.5          # not the same image twice
    want2, ocnt2, _ = oracle_mod.Oracle(s).render_adaptive(2, 1, 10240, sample_base=2 + 10240)
    d2 = np.abs(cnt2.astype(np.int64) - ocnt2.astype(np.int64))
    assert d2.max() <= 1 and (d2 == 0).mean() > 0.9
    same2 = np.kron(d2 == 0, np.ones((32, 32), bool))[:rt.height, :rt.width]
    close2 = np.isclose(second, want2, rtol=2e-3, atol=1e-4).all(axis=-1)       # `second` is a difference of two float films
    assert close2[same2].mean() > 0.98, close2[same2].mean()
    assert np.all(np.abs(second.mean(axis=(0, 1)) / want2.mean(axis=(0, 1)) - 1) < 5e-3)
    rt.clear()
    rt.set_params(primary_reuse=0)
    cnt0, _ = rt.adaptiveRender(2, 1, 10240)
    assert np.array_equal(cnt0, cnt) and np.array_equal(rt.read_film(), img)
    # argument / state errors
    with pytest.raises(rtb.RtbError):
        rt.adaptiveRender(0, 1, 16)
    rt.set_params(scheduler=abi.SCHED_MEGAKERNEL)
    with pytest.raises(rtb.RtbError):
        rt.adaptiveRender(2, 1, 16)
    rt.close()


def test_light_tracer_equals_the_oracle_and_the_reference(rtb, oracle_mod):
    """rtb_render_light = RayTracer::lightTracer (Renderer.h:220-326).  Same uniforms as the oracle: the
    splats land in the same pixels with the same values except where an ulp of the projection crosses a
    pixel border; against the reference's own run (golden block means) statistically."""
    g = np.load(os.path.join(GOLDEN, "cornell256_light_blocks.npz"))
    s = abi.FlatScene.load(os.path.join(GOLDEN, "cornell-box_256.rtbs"))
    rt = rtb.RayTracer(0)
    rt.init(s)
    rt.lightTracer(2)
    img = rt.read_film()
    assert rt.getSPP() == 2
    want, st = oracle_mod.Oracle(s).render_light(2)
    close = np.isclose(img, want, rtol=5e-4, atol=1e-4).all(axis=-1)
    assert close.mean() > 0.99, close.mean()
    assert np.all(np.abs(img.mean(axis=(0, 1)) / want.mean(axis=(0, 1)) - 1) < 5e-3)
    gs = rt.stats()
    assert gs["samples"] == st["paths"]
    assert abs(gs["closest_rays"] / st["closest_rays"] - 1) < 2e-3 and abs(gs["shadow_rays"] / st["shadow_rays"] - 1) < 2e-3
    rt.clear()
    rt.lightTracer(48)
    m = (rt.read_film() / 48).mean(axis=(0, 1))
    assert np.all(np.abs(m / (0.5 * (g["mean_a"] + g["mean_b"])) - 1) < 0.01)
    blocks = raysets.block_mean(rt.read_film() / 48, 16)
    floor48 = np.sqrt(np.mean((g["half_a"] - g["half_b"]) ** 2) / 2)
    rmse = np.sqrt(np.mean((blocks - 0.5 * (g["half_a"] + g["half_b"])) ** 2))
    assert rmse < 3 * floor48 * np.sqrt(1 + 0.5), (rmse, floor48)
    # resumable, and the three traversals give the same bits
    rt.clear()
    rt.lightTracer(1, 0)
    rt.lightTracer(1, 1)
    two = rt.read_film().copy()
    for trav in (abi.TRAV_EXACT, abi.TRAV_WIDE, abi.TRAV_CW, abi.TRAV_Q16):
        rt.set_params(traversal=trav)
        rt.clear()
        rt.lightTracer(2, 0)
        assert np.array_equal(rt.read_film(), two)
    # a scene without area lights adds nothing
    r2 = gpu_scene(rtb, "materialball")
    r2.lightTracer(1)
    assert not r2.read_film().any()
    rt.close()


def test_instant_radiosity_equals_the_oracle_and_the_reference(rtb, oracle_mod):
    """rtb_render_ir = RayTracer::instantRadiosity (Renderer.h:82-218).  Same uniforms and the same VPL
    order as the oracle: per-pixel sums agree to float rounding; the reference's own run statistically."""
    g = np.load(os.path.join(GOLDEN, "cornell256_ir_blocks.npz"))
    s = abi.FlatScene.load(os.path.join(GOLDEN, "cornell-box_256.rtbs"))
    rt = rtb.RayTracer(0)
    rt.init(s)
    rt.instantRadiosity(2)
    img = rt.read_film()
    assert rt.getSPP() == 2
    want, st = oracle_mod.Oracle(s).render_ir(2)
    close = np.isclose(img, want, rtol=5e-4, atol=1e-6).all(axis=-1)
    assert close.mean() > 0.995, close.mean()
    assert np.all(np.abs(img.mean(axis=(0, 1)) / want.mean(axis=(0, 1)) - 1) < 2e-3)
    gs = rt.stats()
    assert abs(gs["closest_rays"] / st["closest_rays"] - 1) < 2e-3 and abs(gs["shadow_rays"] / st["shadow_rays"] - 1) < 2e-3
    rt.clear()
    rt.instantRadiosity(64)
    m = (rt.read_film() / 64).mean(axis=(0, 1))
    halves = np.abs(g["mean_a"] / g["mean_b"] - 1).max()
    assert np.all(np.abs(m / (0.5 * (g["mean_a"] + g["mean_b"])) - 1) < max(0.03, 4 * halves))
    # resumable; the three traversals give the same bits; argument errors
    rt.clear()
    rt.instantRadiosity(1, 0)
    rt.instantRadiosity(1, 1)
    two = rt.read_film().copy()
    # WIDE == FAST bit for bit.  EXACT (the reference's un-culled walk) differs in a few hundred pixels by
    # ~3e-7: a VPL and a shading point on the SAME wall give a visibility ray lying in that wall's plane
    # (cos ~ 1e-8 passes the "<= 0" test), whose plane-intersection t is rounding noise — the one case where
    # the culled trees and the reference's exhaustive walk may accept different triangles.
    for trav in (abi.TRAV_EXACT, abi.TRAV_WIDE, abi.TRAV_CW, abi.TRAV_Q16):
        rt.set_params(traversal=trav)
        rt.clear()
        rt.instantRadiosity(2, 0)
        if trav != abi.TRAV_EXACT:
            assert np.array_equal(rt.read_film(), two)
        else:
            assert np.allclose(rt.read_film(), two, rtol=1e-3, atol=2e-6)
            assert (rt.read_film() != two).any(axis=-1).mean() < 0.02
    with pytest.raises(rtb.RtbError):
        rt.instantRadiosity(1, 0, n_paths=0)
    rt.close()


def test_write_film_is_the_inverse_of_read_film(rtb):
    """rtb_write_film: the hook for RayTracer::denoise (Renderer.h:750-792: film out, external filter,
    film back).  Values round-trip through the 2^-32 fixed-point sums; rendering continues on top."""
    rt = gpu_scene(rtb, "synthetic")
    rt.render(3, 0)
    a = rt.read_film().copy()
    blurred = a.copy()
    blurred[1:-1, 1:-1] = (a[:-2, 1:-1] + a[2:, 1:-1] + a[1:-1, :-2] + a[1:-1, 2:] + a[1:-1, 1:-1]) / 5
    rt.write_film(blurred)
    assert rt.getSPP() == 3
    b = rt.read_film()
    assert np.allclose(b, blurred, rtol=0, atol=2.0 ** -31)
    rt.render(1, 3)
    c = rt.read_film()
    rt.clear()
    rt.render(1, 3)
    assert np.allclose(c - blurred, rt.read_film(), rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        rt.write_film(np.zeros((2, 2, 3), np.float32))


def test_adaptive_render_full_resolution_and_shadow_queue_capacity(rtb, monkeypatch):
    """1024 tiles, tile-major job order: the pool works on a few floor / wall tiles at a time, where
    nearly every vertex queues a shadow ray and the multi-pass shade stage would queue two per slot
    (this overflowed a one-entry-per-slot queue once).  Also the extreme: every path ends at its second
    vertex (max_depth 0) with 7 shade passes."""
    g = np.load(os.path.join(GOLDEN, "cornell_ref_blocks.npz"))
    rt = gpu_scene(rtb, "cornell-box")
    cnt, var = rt.adaptiveRender(2, 1, 10240)
    assert cnt.shape == (32, 32) and cnt.min() >= 1 and cnt.max() > 1000
    img = rt.read_film()
    assert np.all(np.abs(img.mean(axis=(0, 1)) / (0.5 * (g["mean_a"] + g["mean_b"])) - 1) < 0.01)
    from raytracingrenderer_b200 import host_api
    s, _ = host_api.build_soup(1 << 12, 640, 360)
    films = []
    for reuse, passes in ((0, "1"), (1, "7"), (1, "2")):
        monkeypatch.setenv("RTB_PRIMARY_PASSES", passes)
        r2 = rtb.RayTracer(0)
        r2.init(s)
        r2.set_params(max_depth=0, primary_reuse=reuse)
        r2.render(48, 0)
        films.append(r2.read_film().copy())
        r2.close()
    assert np.array_equal(films[0], films[1]) and np.array_equal(films[0], films[2])


def test_adaptive_render_on_a_ragged_image(rtb, oracle_mod):
    """Width and height that are not multiples of 32 (edge tiles, padded pixel slots)."""
    s = synthetic_scene(width=100, height=70)
    rt = rtb.RayTracer(0)
    rt.init(s)
    cnt, var = rt.adaptiveRender(2, 1, 64)
    img = rt.read_film()
    want, ocnt, ovar = oracle_mod.Oracle(s).render_adaptive(2, 1, 64)
    assert cnt.shape == ((rt.height + 31) // 32, (rt.width + 31) // 32)
    assert np.allclose(var, ovar, rtol=1e-3, atol=1e-7)
    d = np.abs(cnt.astype(np.int64) - ocnt.astype(np.int64))
    assert d.max() <= 1
    assert not np.isnan(img).any()
    assert np.all(np.abs(img.mean(axis=(0, 1)) / want.mean(axis=(0, 1)) - 1) < 2e-2)
    rt.close()


@pytest.mark.parametrize("name,spp", [("synthetic", 3), ("cornell-box", 5), ("materialball", 3), ("MaterialsScene", 2)])
def test_primary_hit_table_changes_no_bit_of_the_film(rtb, name, spp, monkeypatch):
    """params.primary_reuse traces each pixel's camera ray once per render call (all samples of a
    pixel share it: pixel centres only, Renderer.h:806-807).  The film must be bit-identical to
    tracing it per sample, for any number of shade passes, and exactly W*H*(spp-1) closest-hit rays
    must disappear from the counters."""
    rt = gpu_scene(rtb, name)
    rt.set_params(primary_reuse=0)
    rt.clear()
    rt.render(spp, 0)
    a = rt.read_film().copy()
    sa = rt.stats()
    for passes in ("1", "2", "7"):
        monkeypatch.setenv("RTB_PRIMARY_PASSES", passes)
        rt2 = rtb.RayTracer(0)
        rt2.init(rt.scene)
        rt2.set_params(primary_reuse=1)
        rt2.render(spp, 0)
        b = rt2.read_film()
        sb = rt2.stats()
        rt2.close()
        assert np.array_equal(a, b), passes
        assert sb["samples"] == sa["samples"]
        assert sb["shadow_rays"] == sa["shadow_rays"]
        assert sa["closest_rays"] - sb["closest_rays"] == rt.width * rt.height * (spp - 1), passes


def test_exact_and_fast_render_identical_films(rtb):
    for name in ("synthetic", "cornell-box"):
        rt = gpu_scene(rtb, name)
        rt.set_params(traversal=abi.TRAV_EXACT)
        rt.render(4, 0)
        a = rt.read_film().copy()
        for trav in (abi.TRAV_FAST, abi.TRAV_WIDE, abi.TRAV_CW, abi.TRAV_Q16):
            rt.set_params(traversal=trav)
            rt.clear()
            rt.render(4, 0)
            assert rt.read_film().tobytes() == a.tobytes(), trav


@pytest.mark.parametrize("name", ["coffee", "bathroom", "materialball_glass", "MaterialsScene"])
def test_exact_fast_wide_films_are_bit_identical_on_the_heavy_scenes(rtb, name):
    """The strongest cheap equivalence gate there is: the fixed-point film is order-independent, so
    film(EXACT) == film(FAST) == film(WIDE) bit for bit means EVERY closest-hit ray of every path (~10^7 per
    scene here; SURVEY F10 found naive culling wrong on 7 of 2.8 M rays) returned the reference traversal's hit
    and every shadow ray its answer — any difference changes a contribution and with it the pixel's bits."""
    rt = gpu_scene(rtb, name)
    spp = 2
    rt.set_params(traversal=abi.TRAV_EXACT, primary_reuse=0)
    rt.render(spp, 0)
    a = rt.read_film().copy()
    sa = rt.stats()
    for trav in (abi.TRAV_FAST, abi.TRAV_WIDE, abi.TRAV_CW, abi.TRAV_Q16):
        rt.set_params(traversal=trav)
        rt.clear()
        rt.render(spp, 0)
        b = rt.read_film()
        sb = rt.stats()
        bad = (a != b).any(axis=-1)
        assert not bad.any(), (trav, int(bad.sum()), np.argwhere(bad)[:5].tolist())
        assert (sa["closest_rays"], sa["shadow_rays"]) == (sb["closest_rays"], sb["shadow_rays"])


@pytest.mark.parametrize("name", ["synthetic", "cornell-box", "coffee"])
def test_device_built_tree_returns_the_same_hits_and_film(rtb, monkeypatch, name):
    """RTB_GPU_BUILD=1: the FAST tree built on the device (linear BVH over the reference's leaves, rtb_gpu_build.cuh — the
    default from 2^20 leaves on) instead of the host's SAH tree.  A different tree over the same leaves: primary hits, the
    recorded ray kinds and the film must not change by a bit, for FAST and for the trees re-encoded from it."""
    host = gpu_scene(rtb, name)
    ids, t, rays = host.primary_hits(abi.TRAV_FAST, want_rays=True)
    host.set_params(primary_reuse=0)
    host.render(2, 0)
    film = host.read_film().copy()
    monkeypatch.setenv("RTB_GPU_BUILD", "1")
    dev = rtb.RayTracer(0)
    dev.init(host.scene)
    ph = dev.trace(rays, traversal=abi.TRAV_EXACT)
    sets = raysets.mixed_set(rays, ph, seed=5, n_each=3000)
    want = host.trace(sets["closest"], traversal=abi.TRAV_EXACT)
    want_any = host.trace(sets["anyhit"], any_hit=True, traversal=abi.TRAV_EXACT)["id"]
    for trav in (abi.TRAV_FAST, abi.TRAV_WIDE, abi.TRAV_CW, abi.TRAV_Q16):
        i2, t2 = dev.primary_hits(trav)
        assert np.array_equal(i2, ids) and t2.tobytes() == t.tobytes(), trav
        assert dev.trace(sets["closest"], traversal=trav).tobytes() == want.tobytes(), trav
        assert np.array_equal(dev.trace(sets["anyhit"], any_hit=True, traversal=trav)["id"], want_any), trav
        dev.set_params(traversal=trav, primary_reuse=0)
        dev.clear()
        dev.render(2, 0)
        assert dev.read_film().tobytes() == film.tobytes(), trav
    dev.close()


def test_megakernel_and_wavefront_schedules_agree(rtb):
    """Same samples, same RNG streams; only the summation order differs."""
    for name in ("synthetic", "cornell-box"):
        rt = gpu_scene(rtb, name)
        rt.set_params(primary_reuse=0)      # the megakernel traces every camera ray
        rt.render(4, 0)
        a = rt.read_film().copy()
        sa = rt.stats()
        rt.set_params(scheduler=abi.SCHED_MEGAKERNEL)
        rt.clear()
        rt.render(4, 0)
        b = rt.read_film()
        sb = rt.stats()
        assert np.allclose(a, b, rtol=1e-5, atol=1e-6)
        for k in ("samples", "closest_rays", "shadow_rays"):
            assert sa[k] == sb[k], k


def test_cornell_image_against_reference_statistics(rtb):
    g = np.load(os.path.join(GOLDEN, "cornell_ref_blocks.npz"))
    rt = gpu_scene(rtb, "cornell-box")
    spp = 256
    rt.render(spp, 0)
    img = rt.read_film() / spp
    ref_mean = 0.5 * (g["mean_a"] + g["mean_b"])
    assert np.all(np.abs(img.mean(axis=(0, 1)) / ref_mean - 1) < 5e-3)       # gate (i): 0.5 %
    blocks = raysets.block_mean(img)
    ref_blocks = 0.5 * (g["half_a"] + g["half_b"])
    floor32 = np.sqrt(np.mean((g["half_a"] - g["half_b"]) ** 2) / 2)
    expect = floor32 * np.sqrt(32) * np.sqrt(1 / spp + 1 / 64)
    rmse = np.sqrt(np.mean((blocks - ref_blocks) ** 2))
    assert rmse < 3 * expect, (rmse, expect)                                  # gate (ii)
    # gate (iv): the reference's committed 144-spp render, 8x8-block RMSE <= 0.003
    assert np.sqrt(np.mean((blocks - g["result_144"]) ** 2)) <= 0.003


@pytest.mark.parametrize("name,spp,rspp", [("cornell-box", 64, 8), ("materialball", 256, 8), ("MaterialsScene", 512, 8), ("coffee", 1024, 8),
                                           ("bathroom", 1024, 2)])
def test_converged_images_at_the_baseline_spp_match_the_reference(rtb, name, spp, rspp):
    """BASELINE.json's configs at their own spp against the UNMODIFIED reference (SURVEY A.7 gates i-iii, the same arithmetic
    as bench.py's per_scene): mean luminance within max(0.5 %, 3 x the difference of two reference runs); 8x8-block RMSE
    between 0.8 and 1.25 x the value the reference's own noise predicts (sigma_1 sqrt(1/N + 1/M), sigma_1 from two
    independent reference half-buffers) — an estimator with any other expectation or variance lands outside; relMSE finite."""
    rs = ref_scene(name)
    rt = gpu_scene(rtb, name)
    a, _, _ = rs.render(rspp, 0, fresh=True)
    b2, _, _ = rs.render(rspp, 0, fresh=False)
    ha, hb = a / rspp, (b2 - a) / rspp
    rt.render(spp, 0)
    img = rt.read_film() / spp
    lum = np.array([0.2126, 0.7152, 0.0722])
    ref = 0.5 * (ha + hb)
    la, lr = float((img @ lum).mean()), float((ref @ lum).mean())
    halves = abs(float(((ha - hb) @ lum).mean())) / lr
    assert abs(la / lr - 1) < max(0.005, 3 * halves), (la, lr, halves)
    ba, bb, bg, br = (raysets.block_mean(x) for x in (ha, hb, img, ref))
    sigma1 = np.sqrt(np.mean((ba - bb) ** 2) / 2 * rspp)
    expect = sigma1 * np.sqrt(1 / spp + 1 / (2 * rspp))
    rmse = np.sqrt(np.mean((bg - br) ** 2))
    assert 0.8 * expect < rmse < 1.25 * expect, (rmse, expect)
    relmse = float(np.mean((img - ref) ** 2 / (ref ** 2 + 1e-2)))
    assert np.isfinite(relmse)


@pytest.mark.parametrize("name,spp", [("MaterialsScene", 64), ("materialball", 64), ("coffee", 32), ("bathroom", 8),
                                      ("materialball_glass", 32), ("materialball_mirror", 32)])
def test_image_statistics_equal_the_reference(rtb, name, spp):
    rs = ref_scene(name)
    rt = gpu_scene(rtb, name)
    rspp = max(2, min(8, spp // 4)) if name != "bathroom" else 2
    a, _, _ = rs.render(rspp, 0, fresh=True)
    b2, _, _ = rs.render(rspp, 0, fresh=False)
    ra, rb = a / rspp, (b2 - a) / rspp
    rt.render(spp, 0)
    img = rt.read_film() / spp
    ref_mean = 0.5 * (ra + rb).mean(axis=(0, 1))
    noise = np.abs((ra - rb).mean(axis=(0, 1))) / ref_mean     # how far two reference runs differ
    assert np.all(np.abs(img.mean(axis=(0, 1)) / ref_mean - 1) < np.maximum(0.01, 3 * noise))
    ba, bb, bo = raysets.block_mean(ra), raysets.block_mean(rb), raysets.block_mean(img)
    floor = np.sqrt(np.mean((ba - bb) ** 2))
    assert np.sqrt(np.mean((bo - 0.5 * (ba + bb)) ** 2)) < 3 * floor


@pytest.mark.parametrize("name", ["synthetic", "materialball", "MaterialsScene", "MaterialsScene_env"])
def test_env_importance_sampler_equals_the_oracle(rtb, oracle_mod, name):
    """RTB_SAMPLING_IMPORTANCE (north_star: "EnvironmentMap importance sampling"; not in the reference, whose
    EnvironmentMap::sample is uniform, Lights.h:143-149): direction, pdf and radiance of the luminance-CDF
    sampler against its restatement in the oracle, <= 1e-5 (radiance: plus the texel slack of a lookup), and
    the density integrates to one: E[1/pdf] over its own samples = 4 pi."""
    rt = gpu_scene(rtb, name)
    env = np.flatnonzero(rt.scene.lights["type"] == abi.LIGHT_ENVMAP)
    assert len(env) == 1
    n = 200000
    rng = np.random.default_rng(17)
    u = rng.random((n, 2), dtype=np.float32)
    u[:64] = np.array([[1e-7, 0.5], [0.9999999, 0.5], [0.5, 1e-7], [0.5, 0.9999999]], np.float32).repeat(16, axis=0)
    li = np.full(n, env[0], np.int32)
    wi = raysets.unit(rng.normal(size=(n, 3)))
    rt.set_params(sampling=abi.SAMPLING_IMPORTANCE)
    g = rt.eval_light(li, wi, u)
    o = oracle_mod.Oracle(rt.scene, sampling=abi.SAMPLING_IMPORTANCE)
    w = o.eval_light(li, wi, u)
    assert rel_err(g["p_or_wi"], w["p_or_wi"]) <= 1e-5
    assert rel_err(g["pdf"], w["pdf"]) <= 1e-5
    assert np.all(g["pdf"] > 0) and np.allclose(np.linalg.norm(g["p_or_wi"], axis=1), 1, atol=1e-5)
    tex = int(rt.scene.lights["tex"][env[0]])
    gdir = np.ascontiguousarray(g["p_or_wi"])
    want = o.eval_light(li, gdir, u)["eval"]               # Light::evaluate at the GPU's own direction
    tol = 1e-5 * np.abs(want) + raysets.env_lookup_slack(rt.scene, tex, gdir, 1e-3) + 1e-12
    assert np.all(np.abs(g["emitted"] - want) <= tol)
    inv = 1.0 / g["pdf"][64:].astype(np.float64)
    assert abs(inv.mean() / (4 * np.pi) - 1) < 4 * inv.std() / np.sqrt(len(inv)) / (4 * np.pi) + 1e-3
    # the sampler prefers bright texels: the mean radiance it returns is above the uniform sampler's
    rt.set_params(sampling=abi.SAMPLING_STRICT)
    s = rt.eval_light(li, wi, u)
    assert rel_err(s["pdf"], np.full(n, 1 / (4 * np.pi), np.float32)) <= 1e-6
    lum = lambda c: c @ np.array([0.2126, 0.7152, 0.0722])
    assert lum(g["emitted"]).mean() >= lum(s["emitted"]).mean()


@pytest.mark.parametrize("name", ["synthetic", "materialball"])
@pytest.mark.parametrize("integ", [abi.INT_PATH, abi.INT_PATH_MIS])
def test_importance_render_equals_the_oracle_sample_for_sample(rtb, oracle_mod, name, integ):
    """pathTrace with computeDirect / computeDirectMIS drawing the env direction from the CDF: same uniforms as
    the oracle, so the films agree sample for sample (both schedules for MIS)."""
    rt = gpu_scene(rtb, name)
    rt.set_params(integrator=integ, sampling=abi.SAMPLING_IMPORTANCE, primary_reuse=0)
    rt.render(2, 0)
    img = rt.read_film().copy()
    want, st = oracle_mod.Oracle(rt.scene, integrator=integ, sampling=abi.SAMPLING_IMPORTANCE).render(2)
    close = np.isclose(img, want, rtol=2e-4, atol=1e-5).all(axis=-1)
    assert close.mean() > 0.995, close.mean()
    assert np.all(np.abs(img.mean(axis=(0, 1)) / want.mean(axis=(0, 1)) - 1) < 5e-3)
    g = rt.stats()
    assert abs(g["shadow_rays"] / st["shadow_rays"] - 1) < 2e-3 and abs(g["closest_rays"] / st["closest_rays"] - 1) < 2e-3
    rt.set_params(scheduler=abi.SCHED_MEGAKERNEL)
    rt.clear()
    rt.render(2, 0)
    assert np.allclose(rt.read_film(), img, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name,spp,rspp", [("MaterialsScene", 128, 8), ("MaterialsScene_env", 128, 8), ("materialball", 64, 8)])
@pytest.mark.parametrize("integ", [abi.INT_PATH, abi.INT_PATH_MIS])
def test_importance_sampling_converges_to_the_reference(rtb, name, spp, rspp, integ):
    """BASELINE config 3 ("MaterialsScene lit by EnvironmentMap (1.hdr) with importance-sampled env lighting +
    MIS"): the env-CDF sampler, with computeDirect and with computeDirectMIS, against the UNMODIFIED reference's
    uniform-sphere render of the same scene.  Same expectation (any positive NEE density; for an
    env-lit scene computeDirectMIS equals computeDirect in expectation, Renderer.h:516-527): mean within
    max(1 %, 3 x the difference of two reference runs), 8x8-block RMSE within 3 x the reference's own noise
    floor — and measurably less noise than the strict sampler at the same spp on the bright map."""
    rs = ref_scene(name)
    rt = gpu_scene(rtb, name)
    a, _, _ = rs.render(rspp, 0, fresh=True)
    b2, _, _ = rs.render(rspp, 0, fresh=False)
    ra, rb = a / rspp, (b2 - a) / rspp
    rt.set_params(integrator=integ, sampling=abi.SAMPLING_IMPORTANCE)
    rt.render(spp, 0)
    img = rt.read_film() / spp
    ref_mean = 0.5 * (ra + rb).mean(axis=(0, 1))
    noise = np.abs((ra - rb).mean(axis=(0, 1))) / ref_mean
    assert np.all(np.abs(img.mean(axis=(0, 1)) / ref_mean - 1) < np.maximum(0.01, 3 * noise))
    ba, bb, bo = raysets.block_mean(ra), raysets.block_mean(rb), raysets.block_mean(img)
    floor = np.sqrt(np.mean((ba - bb) ** 2))
    assert np.sqrt(np.mean((bo - 0.5 * (ba + bb)) ** 2)) < 3 * floor
    if name != "MaterialsScene" and integ == abi.INT_PATH:
        # variance: two independent halves per sampler, block noise of the importance sampler below the strict one's
        def half_noise(sampling):
            rt.set_params(integrator=integ, sampling=sampling)
            out = []
            for begin in (0, 1000):
                rt.clear()
                rt.render(16, begin)
                out.append(raysets.block_mean(rt.read_film() / 16))
            return np.sqrt(np.mean((out[0] - out[1]) ** 2))
        assert half_noise(abi.SAMPLING_IMPORTANCE) < half_noise(abi.SAMPLING_STRICT)


# ------------------------------------------------------------------ film semantics
def test_film_is_a_resumable_deterministic_sum(rtb):
    rt = gpu_scene(rtb, "synthetic")
    rt.render(6, 0)
    a = rt.read_film().copy()
    assert rt.getSPP() == 6
    rt.clear()
    assert rt.getSPP() == 0 and not rt.read_film().any()
    rt.render(2, 0)
    rt.render(4)            # continues at sample 2
    b = rt.read_film().copy()
    assert rt.getSPP() == 6
    assert a.tobytes() == b.tobytes()                   # same samples; fixed-point sums are order-independent
    rt.clear()
    rt.render(6, 0)
    assert rt.read_film().tobytes() == a.tobytes()      # bit-reproducible run to run


def test_render_is_asynchronous_once_the_iteration_count_is_known(rtb, monkeypatch):
    """rtb_render's contract ("asynchronous"): the first render of a configuration probes the device a few times;
    repeating it enqueues the remembered number of iterations and returns while the GPU is still working.  The film is the
    same bits either way, also when the remembered count falls short (a later call enqueues the rest) and when the camera,
    the film or the parameters are touched right after the call."""
    import time
    try:
        rt = gpu_scene(rtb, "coffee")
        spp = 192
    except pytest.skip.Exception:
        rt = gpu_scene(rtb, "cornell-box")
        spp = 96
    rt.set_params(primary_reuse=0)
    rt.render(spp, 0)                        # finds the iteration count (probing path)
    rt.synchronize()
    want = rt.read_film().copy()
    syncs0 = rt.stats()["host_syncs"]
    rt.clear()
    t0 = time.perf_counter()
    rt.render(spp, 0)
    t_call = time.perf_counter() - t0
    rt.synchronize()
    t_all = time.perf_counter() - t0
    assert t_call < 0.5 * t_all, (t_call, t_all)          # the call returned long before the device was done
    assert rt.read_film().tobytes() == want.tobytes()
    assert rt.stats()["host_syncs"] == 1                   # one look at the probe instead of one per batch (clear reset the counter)
    assert syncs0 >= 2
    # touched immediately after the call: clear / camera / parameters / another render all see a completed render
    rt.clear()
    rt.render(spp, 0)
    rt.update_camera(rt.scene.camera)
    rt.set_params(primary_reuse=0)
    rt.render(spp, spp)
    both = rt.read_film().copy()
    rt.clear()
    rt.render(2 * spp, 0)
    assert rt.read_film().tobytes() == both.tobytes()
    # the probing path (0) gives the same bits, and so does a remembered count that falls short (2: half of what is needed
    # is enqueued; the next call on the context finds the pool still alive and enqueues the rest)
    for mode in ("0", "2"):
        monkeypatch.setenv("RTB_ASYNC_RENDER", mode)
        r2 = rtb.RayTracer(0)
        r2.init(rt.scene)
        r2.set_params(primary_reuse=0)
        r2.render(spp, 0)
        r2.clear()
        r2.render(spp, 0)
        assert r2.read_film().tobytes() == want.tobytes(), mode
        st = r2.stats()
        assert st["samples"] == rt.width * rt.height * spp, mode
        r2.clear()
        r2.render(spp, 0)
        r2.render(spp, spp)                  # settles the first one (short in mode 2) before starting
        assert r2.read_film().tobytes() == both.tobytes(), mode
        r2.close()


def test_partitions_compose_to_the_single_gpu_film(rtb):
    """Multi-GPU partitioning, emulated on one device: tile slices are disjoint and compose
    bit-exactly; spp slices compose up to summation order."""
    rt = gpu_scene(rtb, "synthetic")
    rt.render(8, 0)
    full = rt.read_film().copy()
    acc = np.zeros_like(full)
    for r in range(3):
        rt.set_params(partition=abi.PART_TILE, part_rank=r, part_world=3)
        rt.clear()
        rt.render(8, 0)
        part = rt.read_film()
        assert not (acc.astype(bool) & part.astype(bool)).any()
        acc += part
    assert acc.tobytes() == full.tobytes()
    acc = np.zeros_like(full)
    samples = 0
    for r in range(4):
        rt.set_params(partition=abi.PART_SPP, part_rank=r, part_world=4)
        rt.clear()
        rt.render(8, 0)
        acc += rt.read_film()
        samples += rt.stats()["samples"]
    assert samples == rt.width * rt.height * 8
    assert np.allclose(acc, full, rtol=1e-6, atol=1e-7)      # float sum of 4 exact partial films


def test_fixed_point_film_composes_exactly(rtb):
    """The film's master copy is a 64-bit fixed-point sum: adding the int64 buffers of an
    spp-split (what an int64 NCCL reduce does) gives the single-GPU film bit for bit, and the
    film does not depend on the size of the slot pool or on the schedule of the path loop."""
    import torch
    rt = gpu_scene(rtb, "synthetic")
    rt.render(8, 0)
    full = rt.read_film().copy()
    ptr, n = rt.accum_device_ptr()

    class Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}
    acc = torch.as_tensor(Arr(), device="cuda")
    want = acc.clone()
    total = torch.zeros_like(acc)
    for r in range(3):
        rt.set_params(partition=abi.PART_SPP, part_rank=r, part_world=3)
        rt.clear()
        rt.render(8, 0)
        rt.synchronize()
        total += acc
    assert torch.equal(total, want)
    rt.set_params(partition=abi.PART_NONE, part_rank=0, part_world=1)
    rt.clear()
    acc.copy_(total)
    rt.accum_device_ptr()        # marks the float film stale
    rt.set_spp(8)
    assert rt.read_film().tobytes() == full.tobytes() and rt.getSPP() == 8


def test_nccl_reduced_film_equals_the_single_gpu_film(rtb):
    """On hardware, with the real collective: 2 (or 4) ranks under torchrun render their spp / tile slices, the
    int64 accumulators are summed onto rank 0 by NCCL, and that film must equal the 1-GPU film bit for bit
    (tests/tools/nccl_film_check.py).  Needs >= 2 visible GPUs."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "nccl_film_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "NCCL_FILM_OK %d" % world in out.stdout


def test_device_group_of_one_is_a_plain_context(rtb):
    rt = gpu_scene(rtb, "synthetic")
    rt.render(3, 0)
    a = rt.read_film().copy()
    g = rtb.RayTracer([0])
    assert g.group_info() == dict(devices=[0], p2p=[1], gathers_p2p=0, gathers_nccl=0)
    g.init(rt.scene)
    g.render(3, 0)
    assert g.read_film().tobytes() == a.tobytes()
    g.close()
    with pytest.raises(rtb.RtbError):
        rtb.RayTracer([0, 0])
    with pytest.raises(rtb.RtbError):
        rtb.RayTracer([0, 4096])


def test_device_group_film_equals_the_single_gpu_film(rtb, monkeypatch):
    """rtb_create_multi: one context, every GPU of the box (RayTracer::init's numProcs, Renderer.h:52-55).  The
    film after the in-library read-out reduce — the peer-memory kernel, NCCL, and staged copies — must equal the
    1-GPU film bit for bit for sample-sliced and tile-sliced calls, resumed renders, calls with fewer samples than
    devices, and the pass-sharded light tracer / instant radiosity.  Needs >= 2 visible GPUs."""
    n = rtb.lib().rtb_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    devs = list(range(min(n, 4)))
    for name, spp in (("synthetic", 7), ("cornell-box", 5)):
        one = gpu_scene(rtb, name)
        one.render(spp, 0)
        want = one.read_film().copy()
        s1 = one.stats()
        for mode in ("", "nccl", "staged"):
            if mode:
                monkeypatch.setenv("RTB_GROUP_REDUCE", mode)
            else:
                monkeypatch.delenv("RTB_GROUP_REDUCE", raising=False)
            g = rtb.RayTracer(devs)
            info = g.group_info()
            assert info["devices"] == devs
            g.init(one.scene)
            g.render(spp, 0)
            assert g.read_film().tobytes() == want.tobytes(), (name, mode)
            assert g.getSPP() == spp
            sg = g.stats()
            for k in ("samples", "shadow_rays"):
                assert sg[k] == s1[k], (k, mode)
            # every member traces its own primary-hit table (one camera ray per pixel and render call)
            assert sg["closest_rays"] == s1["closest_rays"] + (len(devs) - 1) * one.width * one.height
            info = g.group_info()
            if mode == "nccl":
                assert info["gathers_nccl"] >= 1 and info["gathers_p2p"] == 0
            elif mode == "" and all(info["p2p"]):
                assert info["gathers_p2p"] >= 1 and info["gathers_nccl"] == 0
            # resumed in uneven pieces, a piece with fewer samples than devices (tile split), a read-out in between
            g.clear()
            g.render(1, 0)
            part = g.read_film().copy()
            g.render(spp - 1, 1)
            assert g.read_film().tobytes() == want.tobytes(), (name, mode, "resumed")
            one.clear()
            one.render(1, 0)
            assert part.tobytes() == one.read_film().tobytes()
            one.render(spp - 1, 1)
            # the caller's own partition is refined, not replaced
            g.set_params(partition=abi.PART_TILE, part_rank=0, part_world=1)
            g.clear()
            g.render(spp, 0)
            assert g.read_film().tobytes() == want.tobytes(), (name, mode, "tiles")
            acc = np.zeros_like(want, dtype=np.float64)
            for r in range(3):
                g.set_params(partition=abi.PART_SPP, part_rank=r, part_world=3)
                g.clear()
                g.render(spp, 0)
                acc += g.read_film()
            assert np.allclose(acc, want, rtol=1e-6, atol=1e-7)
            g.close()
    # the light-driven estimators shard their passes
    s = abi.FlatScene.load(os.path.join(GOLDEN, "cornell-box_256.rtbs"))
    monkeypatch.delenv("RTB_GROUP_REDUCE", raising=False)
    one, g = rtb.RayTracer(0), rtb.RayTracer(devs)
    one.init(s)
    g.init(s)
    one.lightTracer(5, 0)
    g.lightTracer(5, 0)
    assert g.read_film().tobytes() == one.read_film().tobytes() and g.getSPP() == 5
    one.clear()
    g.clear()
    one.instantRadiosity(3, 0)
    g.instantRadiosity(3, 0)
    assert g.read_film().tobytes() == one.read_film().tobytes() and g.getSPP() == 3
    # adaptiveRender: every member steers and samples its own tiles; plan and film are the single-GPU ones
    one.clear()
    g.clear()
    c1, v1 = one.adaptiveRender(2, 1, 512)
    cg, vg = g.adaptiveRender(2, 1, 512)
    assert np.array_equal(c1, cg) and np.array_equal(v1, vg)
    assert g.read_film().tobytes() == one.read_film().tobytes() and g.getSPP() == 1
    c1, _ = one.adaptiveRender(2, 1, 512)
    cg, _ = g.adaptiveRender(2, 1, 512)          # the second call draws fresh sample indices on every member
    assert np.array_equal(c1, cg) and g.read_film().tobytes() == one.read_film().tobytes()
    # write_film replaces the whole group's film
    g.write_film(np.ones((256, 256, 3), np.float32))
    assert np.allclose(g.read_film(), 1.0, atol=1e-9)
    one.close()
    g.close()


def test_gaussian_filter_and_tonemap(rtb, oracle_mod):
    rt = gpu_scene(rtb, "synthetic")
    rt.render(4, 0)
    box = rt.read_film().copy()
    rt.set_params(filter=abi.FILTER_GAUSSIAN, filter_radius=2.0, filter_alpha=0.1)
    g = rt.read_film()
    want = oracle_mod.gaussian_splat(box, 2.0, 0.1)
    assert np.allclose(g, want, rtol=1e-5, atol=1e-6)
    rt.set_params(filter=abi.FILTER_BOX)
    t8 = rt.tonemap(1.0).astype(np.int16)
    w8 = oracle_mod.tonemap(box, 4).astype(np.int16)
    assert np.abs(t8 - w8).max() <= 1 and (t8 != w8).mean() < 0.002     # powf ulp at a rounding edge


def test_camera_update(rtb, oracle_mod):
    import refbvh
    rt = gpu_scene(rtb, "synthetic")
    s = rt.scene
    cam = refbvh.look_at_camera([2.0, 1.0, 4.5], [0, 0, 0], [0, 1, 0], 40.0, s.width, s.height)
    rt.update_camera(cam)
    ids, t = rt.primary_hits(abi.TRAV_FAST)
    moved = abi.FlatScene()
    moved.__dict__.update(s.__dict__)
    moved.camera = cam
    oid, ot = oracle_mod.Oracle(moved).primary_hits()
    assert np.array_equal(ids, oid) and t.tobytes() == ot.tobytes()
    rt.update_camera(s.camera)


# ------------------------------------------------------------------ the drop-in boundary itself
def test_reference_host_program_renders_through_the_drop_in_header(rtb, tmp_path):
    """oracle/_ref/dropin_main = the reference's host program shape (Main.cpp) compiled against the
    UNMODIFIED RTBase headers with host/Renderer.h in place of RTBase/Renderer.h: loadScene ->
    RayTracer::init -> render() -> saveHDR on the GPU.  Its film must equal the film the Python
    mirror produces from the same scene, bit for bit."""
    import subprocess
    from oracle import ref
    from raytracingrenderer_b200 import imageio
    exe = os.path.join(ref.REF_DIR, "dropin_main")
    if not os.path.isfile(exe) or not ref.have_scene("cornell-box"):
        pytest.skip("oracle/_ref/dropin_main not built")
    hdr, raw = str(tmp_path / "out.hdr"), str(tmp_path / "film.bin")
    out = subprocess.run([exe, ref.scene_dir("cornell-box"), "8", hdr, raw], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "SPP: 8" in out.stdout
    rt = gpu_scene(rtb, "cornell-box")
    rt.render(8, 0)
    want = rt.read_film()
    got = np.fromfile(raw, "<f4").reshape(want.shape)
    assert got.tobytes() == want.tobytes()
    img = imageio.read_hdr(hdr)                         # Film::save: film / SPP as RGBE
    assert np.all(np.abs(img - want / 8) <= (want / 8).max(axis=-1, keepdims=True) / 100 + 1e-6)
    assert os.path.getsize(hdr + ".png") > 1000          # savePNG after a camera move + clear


def test_standalone_cpp_program_without_any_reference_code(rtb, tmp_path):
    """raytracingrenderer_b200/rtb_render (tools/rtb_render.cpp): Main.cpp's shape built ONLY from this
    repository — host/standalone shims -> rtb_scene.hpp loader + reference-order builder, host/Renderer.h,
    librtb200.so.  Film bit-identical to the Python mirror's; HDR / PNG written by the product's own writers;
    --adaptive and --mis reach the two optional estimators."""
    import subprocess
    from oracle import ref
    from raytracingrenderer_b200 import imageio, build
    exe = build.CLI
    if not os.path.isfile(exe) or not ref.have_scene("cornell-box"):
        pytest.skip("rtb_render or the staged cornell-box scene is missing")
    hdr, raw, png = str(tmp_path / "out.hdr"), str(tmp_path / "film.bin"), str(tmp_path / "out.png")
    out = subprocess.run([exe, ref.scene_dir("cornell-box"), "8", hdr, "--raw", raw, "--png", png], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "SPP 8" in out.stdout
    rt = gpu_scene(rtb, "cornell-box")
    rt.render(8, 0)
    want = rt.read_film()
    got = np.fromfile(raw, "<f4").reshape(want.shape)
    assert got.tobytes() == want.tobytes()
    img = imageio.read_hdr(hdr)
    assert np.all(np.abs(img - want / 8) <= (want / 8).max(axis=-1, keepdims=True) / 100 + 1e-6)
    assert open(png, "rb").read(8) == b"\x89PNG\r\n\x1a\n" and os.path.getsize(png) > 1000
    g = np.load(os.path.join(GOLDEN, "cornell_mis_blocks.npz"))
    out = subprocess.run([exe, ref.scene_dir("cornell-box"), "16", hdr, "--mis", "--raw", raw], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    m = (np.fromfile(raw, "<f4").reshape(want.shape) / 16).mean(axis=(0, 1))
    assert np.all(np.abs(m / (0.5 * (g["mean_a"] + g["mean_b"])) - 1) < 0.01)
    out = subprocess.run([exe, ref.scene_dir("cornell-box"), "2", hdr, "--adaptive", "--raw", raw], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "SPP 2" in out.stdout
    plain = np.load(os.path.join(GOLDEN, "cornell_ref_blocks.npz"))
    m = (np.fromfile(raw, "<f4").reshape(want.shape) / 2).mean(axis=(0, 1))
    assert np.all(np.abs(m / (0.5 * (plain["mean_a"] + plain["mean_b"])) - 1) < 0.01)
    for flag in ("--light", "--ir"):
        out = subprocess.run([exe, ref.scene_dir("cornell-box"), "2", hdr, flag, "--raw", raw], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        assert "SPP 2" in out.stdout
        f = np.fromfile(raw, "<f4").reshape(want.shape)
        assert np.isfinite(f).all() and f.mean() > 0.01, flag
    # a scene directory that does not exist is an error message and a non-zero exit, not a crash
    out = subprocess.run([exe, str(tmp_path / "nowhere"), "1", hdr], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0


def test_soup_config_primary_plus_one_bounce(rtb, oracle_mod):
    """SURVEY 8d cfg 5 at test size: random-triangle soup built by the product's host layer
    (reference-order BVH), max_depth 0 = primary + one diffuse bounce, white background light."""
    from raytracingrenderer_b200 import host_api
    s, _ = host_api.build_soup(1 << 14, 320, 180)
    rt = rtb.RayTracer(0)
    rt.init(s)
    rt.set_params(max_depth=0, primary_reuse=0)
    o = oracle_mod.Oracle(s, max_depth=0)
    for trav in TRAVS:
        ids, t = rt.primary_hits(trav)
        oid, ot = o.primary_hits()
        assert np.array_equal(ids, oid) and t.tobytes() == ot.tobytes()
    rt.clear()              # the parity entry points above count rays too
    rt.render(4, 0)
    img = rt.read_film()
    want, st = o.render(4)
    close = np.isclose(img, want, rtol=2e-4, atol=1e-5).all(axis=-1)
    assert close.mean() > 0.995
    g = rt.stats()
    assert g["closest_rays"] <= 2 * g["samples"] and g["shadow_rays"] <= 2 * g["samples"]
    assert abs(g["closest_rays"] / st["closest_rays"] - 1) < 2e-3
    rt.close()


# ------------------------------------------------------------------ edge cases
def _tiny_scene(n_tris, width, height, background=(0.25, 0.5, 1.0)):
    """0, 1 or a few triangles in front of the camera, constant background light, odd film size."""
    import refbvh
    F = np.float32
    mats, texs, texels = refbvh.standard_materials()
    v0 = np.array([[-1, -1, 0], [0.2, -0.8, 0.5], [-0.9, 0.1, -0.4]], F)[:n_tris]
    v1 = np.array([[1, -1, 0], [0.9, -0.7, 0.6], [-0.2, 0.2, -0.5]], F)[:n_tris]
    v2 = np.array([[0, 1, 0], [0.5, 0.4, 0.4], [-0.6, 0.9, -0.3]], F)[:n_tris]
    nrm = np.tile(np.array([[0, 0, 1]], F), (n_tris, 1))
    uv = np.zeros((n_tris, 3, 2), F)
    ti, ts = refbvh._tri_records(v0, v1, v2, nrm, nrm, nrm, uv, np.zeros(n_tris, np.uint32))
    cam = refbvh.look_at_camera([0.1, 0.2, 4.0], [0, 0, 0], [0, 1, 0], 45.0, width, height)
    return refbvh.assemble(ti, ts, mats, texs, texels, cam, None, background)


@pytest.mark.parametrize("n_tris", [0, 1, 3])
@pytest.mark.parametrize("size", [(13, 7), (33, 5)])
def test_degenerate_scenes_and_ragged_film_sizes(rtb, oracle_mod, n_tris, size):
    """Empty scene, single-leaf scene (the root is a leaf), film sizes that are not multiples of the
    8x4 pixel tiles, every traversal and both schedules."""
    s = _tiny_scene(n_tris, *size)
    rt = rtb.RayTracer(0)
    rt.init(s)
    o = oracle_mod.Oracle(s)
    oid, ot = o.primary_hits()
    want, st = o.render(3)
    for trav in TRAVS:
        ids, t = rt.primary_hits(trav)
        assert np.array_equal(ids, oid) and t.tobytes() == ot.tobytes()
        for sched in (abi.SCHED_WAVEFRONT, abi.SCHED_MEGAKERNEL):
            rt.set_params(traversal=trav, scheduler=sched)
            rt.clear()
            rt.render(3, 0)
            img = rt.read_film()
            assert img.shape == (size[1], size[0], 3)
            assert np.allclose(img, want, rtol=2e-4, atol=1e-5), (trav, sched)
            assert rt.stats()["samples"] == size[0] * size[1] * 3
    if n_tris == 0:
        assert np.allclose(want / 3, np.array([0.25, 0.5, 1.0], np.float32))     # every ray sees the background
    rt.close()


def test_more_ranks_than_samples_and_late_sample_indices(rtb, oracle_mod):
    """spp partition with world > spp (some ranks get nothing) and a render that resumes at a large
    sample index: the RNG is keyed by the global sample index."""
    rt = gpu_scene(rtb, "synthetic")
    rt.render(2, 1000)
    full = rt.read_film().copy()
    want, _ = oracle_mod.Oracle(rt.scene).render(2, spp_begin=1000)
    assert np.isclose(full, want, rtol=2e-4, atol=1e-5).all(axis=-1).mean() > 0.995
    acc = np.zeros_like(full)
    empty = 0
    for r in range(5):
        rt.set_params(partition=abi.PART_SPP, part_rank=r, part_world=5)
        rt.clear()
        rt.render(2, 1000)
        part = rt.read_film()
        empty += int(not part.any())
        acc += part
    assert empty == 3                                       # samples 1000, 1001 -> ranks 0 and 1
    assert np.allclose(acc, full, rtol=1e-6, atol=1e-7)


def test_every_mode_combination_runs(rtb):
    """tests/tools/sanitize_case.py as a test: all traversals x integrators x schedules x partitions."""
    import runpy
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "sanitize_case.py"), run_name="__main__")
