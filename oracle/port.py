"""TEST INFRASTRUCTURE (oracle/): ctypes front end of the plain-C restatement
(oracle/rtb_oracle.c).  Same call surface as raytracingrenderer_b200.RayTracer's parity
entry points so tests read `gpu.x(...)` vs `oracle.x(...)`.  Import only from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from raytracingrenderer_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rtb_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "librtb_oracle.so")

_lib = None


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC, os.path.join(os.path.dirname(HERE), "include", "rtb.h")]
    if (not force and os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps)):
        return LIB
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-std=c11", "-D_GNU_SOURCE", "-shared", "-fPIC",
                           SRC, "-o", LIB, "-lm", "-lpthread"])
    return LIB


def _p(a):
    return None if a is None else a.ctypes.data


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
        L.oracle_render.argtypes = [vp, vp, u32, u32, i32, vp, vp]
        L.oracle_render_counts.argtypes = [vp, vp, u32, u32, i32, vp, vp]
        L.oracle_render_adaptive.argtypes = [vp, vp, u32, u32, u32, i32, vp, vp, vp]
        L.oracle_render_adaptive_at.argtypes = [vp, vp, u32, u32, u32, u32, i32, vp, vp, vp]
        L.oracle_render_light.argtypes = [vp, vp, u32, u32, vp, vp]
        L.oracle_render_ir.argtypes = [vp, vp, u32, u32, u32, i32, vp, vp]
        L.oracle_camera_derive.argtypes = [vp, vp]
        L.oracle_primary_hits.argtypes = [vp, f32, vp, vp, vp]
        L.oracle_trace.argtypes = [vp, f32, i32, vp, u64, vp]
        L.oracle_visible.argtypes = [vp, f32, vp, u64, vp]
        L.oracle_shading_data.argtypes = [vp, vp, vp, u64, vp]
        L.oracle_eval_bsdf.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp, vp]
        L.oracle_eval_light.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp]
        L.oracle_eval_light_p.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, vp, vp]
        L.oracle_rng_draws.argtypes = [u32, u32, u32, u32, vp]
        L.oracle_tonemap.argtypes = [vp, u32, i32, f32, vp]
        L.oracle_gaussian_splat.argtypes = [i32, i32, f32, f32, vp, vp]
        _lib = L
    return _lib


def default_params():
    """The reference's constants (Renderer.h:20, Geometry.h:60, Renderer.h:353)."""
    p = abi.Params()
    p.max_depth, p.epsilon, p.rr_cap = 4, 1e-4, 0.9
    p.integrator, p.sampling, p.traversal, p.filter = abi.INT_PATH, abi.SAMPLING_STRICT, abi.TRAV_EXACT, abi.FILTER_BOX
    p.filter_radius, p.filter_alpha, p.seed = 2.0, 0.1, 1
    p.partition, p.part_rank, p.part_world, p.cull_rel = abi.PART_NONE, 0, 1, 1e-5
    return p


def rng_draws(seed, pixel, sample, n):
    out = np.zeros(n, "<f4")
    lib().oracle_rng_draws(int(seed), int(pixel), int(sample), int(n), _p(out))
    return out


def tonemap(film_sum, spp, exposure=1.0):
    f = np.ascontiguousarray(film_sum, "<f4")
    out = np.zeros(f.shape, np.uint8)
    lib().oracle_tonemap(_p(f), f.size // 3, int(spp), float(exposure), _p(out))
    return out


def gaussian_splat(colours, radius=2.0, alpha=0.1):
    col = np.ascontiguousarray(colours, "<f4")
    h, w, _ = col.shape
    out = np.zeros_like(col)
    lib().oracle_gaussian_splat(w, h, float(radius), float(alpha), _p(col), _p(out))
    return out


class Oracle:
    """CPU restatement bound to one flat scene."""

    def __init__(self, scene, **params):
        self.L = lib()
        self.scene = scene
        self.desc = scene.desc()
        self.width, self.height = scene.width, scene.height
        self.params = default_params()
        self.set_params(**params)

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise AttributeError(k)
            setattr(self.params, k, v)

    def render(self, spp, spp_begin=0, threads=None, film=None):
        """-> (film_sum [H,W,3], stats dict)"""
        if film is None:
            film = np.zeros((self.height, self.width, 3), "<f4")
        threads = threads or os.cpu_count() or 1
        st = np.zeros(3, np.uint64)
        self.L.oracle_render(C.addressof(self.desc), C.addressof(self.params), int(spp_begin), int(spp), int(threads),
                             _p(film), _p(st))
        return film, dict(samples=int(st[0]), closest_rays=int(st[1]), shadow_rays=int(st[2]))

    def render_adaptive(self, init_samples=2, min_samples=1, max_samples=10240, threads=None, film=None, sample_base=0):
        """RayTracer::adaptiveRender -> (film_sum + one mean image, tile_samples, tile_variance)"""
        if film is None:
            film = np.zeros((self.height, self.width, 3), "<f4")
        ty, tx = (self.height + 31) // 32, (self.width + 31) // 32
        samples, var = np.zeros((ty, tx), np.uint32), np.zeros((ty, tx), np.float32)
        self.L.oracle_render_adaptive_at(C.addressof(self.desc), C.addressof(self.params), int(sample_base), int(init_samples),
                                         int(min_samples), int(max_samples), int(threads or os.cpu_count() or 1), _p(film),
                                         _p(samples), _p(var))
        return film, samples, var

    def render_light(self, passes, pass_begin=0, film=None):
        """RayTracer::lightTracer x passes -> (film_sum, stats)."""
        if film is None:
            film = np.zeros((self.height, self.width, 3), "<f4")
        st = np.zeros(3, np.uint64)
        self.L.oracle_render_light(C.addressof(self.desc), C.addressof(self.params), int(pass_begin), int(passes), _p(film), _p(st))
        return film, dict(paths=int(st[0]), closest_rays=int(st[1]), shadow_rays=int(st[2]))

    def render_ir(self, passes, pass_begin=0, n_paths=50, threads=None, film=None):
        """RayTracer::instantRadiosity x passes -> (film_sum, stats)."""
        if film is None:
            film = np.zeros((self.height, self.width, 3), "<f4")
        st = np.zeros(4, np.uint64)
        self.L.oracle_render_ir(C.addressof(self.desc), C.addressof(self.params), int(pass_begin), int(passes), int(n_paths),
                                int(threads or os.cpu_count() or 1), _p(film), _p(st))
        return film, dict(pixels=int(st[0]), closest_rays=int(st[1]), shadow_rays=int(st[2]), vpls=int(st[3]))

    def camera_ext(self):
        """include/rtb.h: rtb_camera_derive -> float32[36] = proj, world_to_cam, view_dir, afilm."""
        out = np.zeros(36, "<f4")
        if not self.L.oracle_camera_derive(C.addressof(self.desc.camera), _p(out)):
            raise ValueError("singular camera matrices")
        return out

    def render_counts(self, spp, spp_begin=0, threads=None):
        """render() + the CANONICAL traversal work of every ray it traced."""
        film = np.zeros((self.height, self.width, 3), "<f4")
        threads = threads or os.cpu_count() or 1
        st = np.zeros(11, np.uint64)
        self.L.oracle_render_counts(C.addressof(self.desc), C.addressof(self.params), int(spp_begin), int(spp),
                                    int(threads), _p(film), _p(st))
        keys = ("samples", "closest_rays", "shadow_rays", "closest_box", "closest_tri", "shadow_box", "shadow_tri",
                "closest_box_max", "shadow_box_max", "closest_rays_over_20k_boxes", "nan_rays")
        return film, {k: int(v) for k, v in zip(keys, st)}

    def primary_hits(self, want_rays=False):
        n = self.width * self.height
        ids, t = np.zeros(n, "<u4"), np.zeros(n, "<f4")
        rays = np.zeros(n, abi.ray_dt) if want_rays else None
        self.L.oracle_primary_hits(C.addressof(self.desc), self.params.epsilon, _p(ids), _p(t), _p(rays))
        return (ids, t, rays) if want_rays else (ids, t)

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, abi.ray_dt)
        hits = np.zeros(len(rays), abi.hit_dt)
        self.L.oracle_trace(C.addressof(self.desc), self.params.epsilon, 1 if any_hit else 0, _p(rays), len(rays),
                            _p(hits))
        return hits

    def visible(self, p1p2):
        p = np.ascontiguousarray(p1p2, "<f4").reshape(-1, 6)
        out = np.zeros(len(p), np.uint8)
        self.L.oracle_visible(C.addressof(self.desc), self.params.epsilon, _p(p), len(p), _p(out))
        return out

    def shading_data(self, rays, hits):
        rays = np.ascontiguousarray(rays, abi.ray_dt)
        hits = np.ascontiguousarray(hits, abi.hit_dt)
        out = np.zeros(len(rays), abi.shading_dt)
        self.L.oracle_shading_data(C.addressof(self.desc), _p(rays), _p(hits), len(rays), _p(out))
        return out

    def eval_bsdf(self, sd, wi, u):
        sd = np.ascontiguousarray(sd, abi.shading_dt)
        n = len(sd)
        wi = np.ascontiguousarray(wi, "<f4").reshape(n, 3)
        u = np.ascontiguousarray(u, "<f4").reshape(n, 3)
        out = dict(eval=np.zeros((n, 3), "<f4"), pdf=np.zeros(n, "<f4"), s_wi=np.zeros((n, 3), "<f4"),
                   s_f=np.zeros((n, 3), "<f4"), s_pdf=np.zeros(n, "<f4"))
        rc = self.L.oracle_eval_bsdf(C.addressof(self.desc), _p(sd), _p(wi), _p(u), n, _p(out["eval"]),
                                     _p(out["pdf"]), _p(out["s_wi"]), _p(out["s_f"]), _p(out["s_pdf"]))
        if rc != 0:
            raise ValueError("oracle_eval_bsdf: bad material index")
        return out

    def eval_light(self, light, wi, u):
        light = np.ascontiguousarray(light, "<i4")
        n = len(light)
        wi = np.ascontiguousarray(wi, "<f4").reshape(n, 3)
        u = np.ascontiguousarray(u, "<f4").reshape(n, 2)
        out = dict(p_or_wi=np.zeros((n, 3), "<f4"), emitted=np.zeros((n, 3), "<f4"), pdf=np.zeros(n, "<f4"),
                   eval=np.zeros((n, 3), "<f4"))
        rc = self.L.oracle_eval_light_p(C.addressof(self.desc), C.addressof(self.params), _p(light), _p(wi), _p(u), n,
                                        _p(out["p_or_wi"]), _p(out["emitted"]), _p(out["pdf"]), _p(out["eval"]))
        if rc != 0:
            raise ValueError("oracle_eval_light: bad light index")
        return out
