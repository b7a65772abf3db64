"""TEST INFRASTRUCTURE (oracle/): ctypes front end of oracle/_ref/librtref*.so — the
UNMODIFIED reference renderer compiled by oracle/build_ref.py.  Import only from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os

import numpy as np

from raytracingrenderer_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SCENES_DIR = os.path.join(os.path.dirname(HERE), "scenes", "_staged")

_libs = {}


def available(variant=""):
    return os.path.isfile(os.path.join(REF_DIR, "librtref%s.so" % variant))


def scene_dir(name):
    return os.path.join(SCENES_DIR, name)


def have_scene(name):
    return os.path.isfile(os.path.join(scene_dir(name), "scene.json"))


def _lib(variant=""):
    if variant not in _libs:
        lib = C.CDLL(os.path.join(REF_DIR, "librtref%s.so" % variant))
        lib.ref_load_scene.restype = C.c_void_p
        lib.ref_load_scene.argtypes = [C.c_char_p]
        vp = C.c_void_p
        lib.ref_info.argtypes = [vp, vp]
        lib.ref_arg_order.argtypes = [vp]
        lib.ref_flatten.argtypes = [vp, C.c_char_p]
        lib.ref_primary_hits.argtypes = [vp, vp, vp, vp]
        lib.ref_trace.argtypes = [vp, C.c_int, vp, C.c_uint64, vp]
        lib.ref_visible.argtypes = [vp, vp, C.c_uint64, vp]
        lib.ref_shading_data.argtypes = [vp, vp, vp, C.c_uint64, vp]
        lib.ref_eval_bsdf.argtypes = [vp, vp, vp, vp, C.c_uint64, vp, vp, vp, vp, vp]
        lib.ref_eval_light.argtypes = [vp, vp, vp, vp, C.c_uint64, vp, vp, vp, vp]
        lib.ref_eval_background.argtypes = [vp, vp, C.c_uint64, vp]
        lib.ref_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
        lib.ref_render_adaptive.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp]
        lib.ref_render_light.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        lib.ref_render_ir.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        lib.ref_camera_ext.argtypes = [vp, vp]
        lib.ref_decode_hdr.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), vp, C.c_uint64]
        lib.ref_decode_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), vp, C.c_uint64]
        lib.ref_aov.argtypes = [vp, C.c_int, vp]
        lib.ref_tonemap.argtypes = [vp, vp, C.c_int, C.c_float, vp]
        lib.ref_gaussian_splat.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, vp, vp]
        _libs[variant] = lib
    return _libs[variant]


def _p(a):
    return None if a is None else a.ctypes.data


def gaussian_splat(colours, radius=2.0, alpha=0.1):
    h, w, _ = colours.shape
    col = np.ascontiguousarray(colours, "<f4")
    out = np.zeros((h, w, 3), "<f4")
    _lib().ref_gaussian_splat(w, h, radius, alpha, _p(col), _p(out))
    return out


class RefScene:
    """A scene loaded by the reference's own loadScene (RTBase/SceneLoader.h:237)."""

    def __init__(self, name_or_dir, variant=""):
        d = name_or_dir if os.path.isdir(name_or_dir) else scene_dir(name_or_dir)
        if not os.path.isfile(os.path.join(d, "scene.json")):
            raise FileNotFoundError(d)
        self.lib = _lib(variant)
        self.dir = d
        self.h = self.lib.ref_load_scene(d.encode())
        info = np.zeros(6, np.int32)
        self.lib.ref_info(self.h, _p(info))
        self.width, self.height, self.n_tris, self.n_materials, self.n_lights, self.hw_threads = map(int, info)

    def flatten(self, path):
        if self.lib.ref_flatten(self.h, path.encode()) != 0:
            raise IOError("ref_flatten -> %s failed" % path)
        return abi.FlatScene.load(path)

    def primary_hits(self, want_rays=False):
        n = self.width * self.height
        ids = np.zeros(n, "<u4")
        t = np.zeros(n, "<f4")
        rays = np.zeros(n, abi.ray_dt) if want_rays else None
        self.lib.ref_primary_hits(self.h, _p(ids), _p(t), _p(rays))
        return (ids, t, rays) if want_rays else (ids, t)

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, abi.ray_dt)
        hits = np.zeros(len(rays), abi.hit_dt)
        self.lib.ref_trace(self.h, 1 if any_hit else 0, _p(rays), len(rays), _p(hits))
        return hits

    def visible(self, p1p2):
        p = np.ascontiguousarray(p1p2, "<f4").reshape(-1, 6)
        out = np.zeros(len(p), np.uint8)
        self.lib.ref_visible(self.h, _p(p), len(p), _p(out))
        return out

    def shading_data(self, rays, hits):
        rays = np.ascontiguousarray(rays, abi.ray_dt)
        hits = np.ascontiguousarray(hits, abi.hit_dt)
        out = np.zeros(len(rays), abi.shading_dt)
        self.lib.ref_shading_data(self.h, _p(rays), _p(hits), len(rays), _p(out))
        return out

    def eval_bsdf(self, sd, wi, u):
        sd = np.ascontiguousarray(sd, abi.shading_dt)
        n = len(sd)
        wi = np.ascontiguousarray(wi, "<f4").reshape(n, 3)
        u = np.ascontiguousarray(u, "<f4").reshape(n, 3)
        out = dict(eval=np.zeros((n, 3), "<f4"), pdf=np.zeros(n, "<f4"), s_wi=np.zeros((n, 3), "<f4"),
                   s_f=np.zeros((n, 3), "<f4"), s_pdf=np.zeros(n, "<f4"))
        rc = self.lib.ref_eval_bsdf(self.h, _p(sd), _p(wi), _p(u), n, _p(out["eval"]), _p(out["pdf"]),
                                    _p(out["s_wi"]), _p(out["s_f"]), _p(out["s_pdf"]))
        if rc != 0:
            raise ValueError("ref_eval_bsdf: bad material index")
        return out

    def eval_light(self, light, wi, u):
        light = np.ascontiguousarray(light, "<i4")
        n = len(light)
        wi = np.ascontiguousarray(wi, "<f4").reshape(n, 3)
        u = np.ascontiguousarray(u, "<f4").reshape(n, 2)
        out = dict(p_or_wi=np.zeros((n, 3), "<f4"), emitted=np.zeros((n, 3), "<f4"), pdf=np.zeros(n, "<f4"),
                   eval=np.zeros((n, 3), "<f4"))
        rc = self.lib.ref_eval_light(self.h, _p(light), _p(wi), _p(u), n, _p(out["p_or_wi"]), _p(out["emitted"]),
                                     _p(out["pdf"]), _p(out["eval"]))
        if rc != 0:
            raise ValueError("ref_eval_light: bad light index")
        return out

    def eval_background(self, wi):
        wi = np.ascontiguousarray(wi, "<f4").reshape(-1, 3)
        out = np.zeros_like(wi)
        self.lib.ref_eval_background(self.h, _p(wi), len(wi), _p(out))
        return out

    def render(self, spp, threads=0, fresh=True):
        """-> (film_sum[H,W,3], spp_total, seconds).  RayTracer::render() x spp."""
        film = np.zeros((self.height, self.width, 3), "<f4")
        secs = C.c_double(0)
        total = self.lib.ref_render(self.h, int(spp), int(threads), 1 if fresh else 0, _p(film), C.addressof(secs))
        return film, total, secs.value

    def render_adaptive(self, threads=0, fresh=True):
        """One render() with adaptiveRender switched on -> (film_sum, tile_samples, tile_variance, seconds)."""
        film = np.zeros((self.height, self.width, 3), "<f4")
        ty, tx = (self.height + 31) // 32, (self.width + 31) // 32
        var, cnt = np.zeros((ty, tx), "<f4"), np.zeros((ty, tx), np.int32)
        secs = C.c_double(0)
        self.lib.ref_render_adaptive(self.h, int(threads), 1 if fresh else 0, _p(film), _p(var), _p(cnt), C.addressof(secs))
        return film, cnt, var, secs.value

    def render_light(self, passes, fresh=True):
        """passes x lightTracer() -> (film_sum, seconds)."""
        film = np.zeros((self.height, self.width, 3), "<f4")
        secs = C.c_double(0)
        self.lib.ref_render_light(self.h, int(passes), 1 if fresh else 0, _p(film), C.addressof(secs))
        return film, secs.value

    def render_ir(self, passes, threads=0, fresh=True):
        """passes x instantRadiosity() -> (film_sum, seconds, total VPLs)."""
        film = np.zeros((self.height, self.width, 3), "<f4")
        secs = C.c_double(0)
        nv = C.c_uint64(0)
        self.lib.ref_render_ir(self.h, int(passes), int(threads), 1 if fresh else 0, _p(film), C.addressof(secs), C.addressof(nv))
        return film, secs.value, nv.value

    def camera_ext(self):
        """Camera::projectionMatrix, cameraToView, viewDirection, Afilm as one float32[36]."""
        out = np.zeros(36, "<f4")
        self.lib.ref_camera_ext(self.h, _p(out))
        return out

    def aov(self, kind):
        k = {"albedo": 0, "normals": 1, "direct": 2}[kind]
        out = np.zeros((self.height, self.width, 3), "<f4")
        self.lib.ref_aov(self.h, k, _p(out))
        return out

    def tonemap(self, film_sum, spp, exposure=1.0):
        f = np.ascontiguousarray(film_sum, "<f4")
        out = np.zeros((self.height, self.width, 3), np.uint8)
        self.lib.ref_tonemap(self.h, _p(f), int(spp), float(exposure), _p(out))
        return out


def decode_image(path):
    """stbi_load of the reference on one file -> uint8 [H, W, channels]."""
    lib = _lib("")
    w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
    if lib.ref_decode_image(path.encode(), C.byref(w), C.byref(h), C.byref(c), None, 0) != 0:
        raise RuntimeError("stbi_load failed on " + path)
    out = np.zeros((h.value, w.value, c.value), np.uint8)
    lib.ref_decode_image(path.encode(), C.byref(w), C.byref(h), C.byref(c), _p(out), out.size)
    return out


def decode_hdr(path):
    """stbi_loadf of the reference on one .hdr file -> float32 [H, W, channels]."""
    lib = _lib("")
    w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
    if lib.ref_decode_hdr(path.encode(), C.byref(w), C.byref(h), C.byref(c), None, 0) != 0:
        raise RuntimeError("stbi_loadf failed on " + path)
    out = np.zeros((h.value, w.value, c.value), "<f4")
    lib.ref_decode_hdr(path.encode(), C.byref(w), C.byref(h), C.byref(c), _p(out), out.size)
    return out
