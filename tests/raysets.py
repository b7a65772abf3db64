"""Seeded ray sets for the traversal-equivalence tests: primary rays, random
bounce rays leaving the primary hit points, shadow segments toward random scene points and
axis-aligned rays (a direction component exactly 0 — the 0*inf = NaN slab case the reference
treats specially)."""
import numpy as np

from raytracingrenderer_b200 import abi

EPS = np.float32(1e-4)


def hit_points(rays, hits):
    m = hits["id"] != abi.MISS_ID
    x = rays["o"][m] + rays["d"][m] * hits["t"][m][:, None]
    return x.astype(np.float32), m


def unit(v):
    v = np.asarray(v, np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def make_rays(o, d, tmax=None):
    r = np.zeros(len(o), abi.ray_dt)
    r["o"], r["d"] = o, d
    r["tmax"] = abi.FLT_MAX if tmax is None else tmax
    return r


def bounce_rays(points, rng, n):
    idx = rng.integers(0, len(points), n)
    d = unit(rng.normal(size=(n, 3)))
    o = points[idx] + d * EPS
    return make_rays(o.astype(np.float32), d)


def axis_rays(points, rng, n):
    idx = rng.integers(0, len(points), n)
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                     [1, 1, 0], [0, 1, -1], [-1, 0, 1]], np.float32)
    d = axes[rng.integers(0, len(axes), n)]
    d = unit(d)
    return make_rays(points[idx].astype(np.float32), d)


def shadow_segments(points, rng, n):
    """p1p2 pairs for Scene::visible."""
    a = points[rng.integers(0, len(points), n)]
    b = points[rng.integers(0, len(points), n)]
    keep = np.linalg.norm(a - b, axis=1) > 1e-3
    return np.concatenate([a[keep], b[keep]], axis=1).astype(np.float32)


def anyhit_rays(points, rng, n):
    idx = rng.integers(0, len(points), n)
    d = unit(rng.normal(size=(n, 3)))
    tmax = rng.uniform(0.05, 5.0, n).astype(np.float32)
    return make_rays((points[idx] + d * EPS).astype(np.float32), d, tmax)


def mixed_set(primary_rays, primary_hits, seed, n_each):
    """-> dict(closest=rays, anyhit=rays with tmax, segments=p1p2)"""
    rng = np.random.default_rng(seed)
    pts, _ = hit_points(primary_rays, primary_hits)
    sub = primary_rays[rng.integers(0, len(primary_rays), n_each)]
    closest = np.concatenate([sub, bounce_rays(pts, rng, n_each), axis_rays(pts, rng, n_each)])
    return dict(closest=closest, anyhit=np.concatenate([anyhit_rays(pts, rng, n_each), axis_rays(pts, rng, n_each // 4)]),
                segments=shadow_segments(pts, rng, n_each))


def block_mean(img, b=8):
    h, w, c = img.shape
    hh, ww = (h // b) * b, (w // b) * b
    return img[:hh, :ww].reshape(hh // b, b, ww // b, b, c).mean(axis=(1, 3))


def env_lookup_slack(scene, tex, dirs, delta_texels):
    """Per-direction, per-channel bound on how much EnvironmentMap::evaluate (Lights.h:158-165 +
    Imaging.h:72-94) can change when the texel coordinate moves by `delta_texels`: delta x the
    value range of the 3x3 texels around the lookup."""
    T = scene.textures[tex]
    W, H = int(T["width"]), int(T["height"])
    px = scene.texels.reshape(-1, 3)[int(T["offset"]):int(T["offset"]) + W * H].reshape(H, W, 3)
    d = np.asarray(dirs, np.float64)
    u = np.arctan2(d[:, 2], d[:, 0])
    u = np.where(u < 0, u + 2 * np.pi, u) / (2 * np.pi)
    v = np.arccos(np.clip(d[:, 1], -1, 1)) / np.pi
    x, y = np.floor(u * W).astype(int), np.floor(v * H).astype(int)
    lo = np.full((len(d), 3), np.inf)
    hi = np.full((len(d), 3), -np.inf)
    for dy in (-1, 0, 1, 2):
        for dx in (-1, 0, 1, 2):
            t = px[(y + dy) % H, (x + dx) % W]
            lo, hi = np.minimum(lo, t), np.maximum(hi, t)
    return delta_texels * (hi - lo)
