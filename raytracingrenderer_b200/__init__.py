"""raytracingrenderer_b200 — B200 (sm_100a) implementation of RTBase's per-pixel path-tracing
loop behind the reference's scene / renderer API.

The product is the C-ABI shared library ``librtb200.so`` (include/rtb.h) built from
``csrc/`` by ``build.py``.  This module is the thin Python host mirror used by bench.py and
the tests: ``RayTracer`` has the reference's ``RayTracer`` surface (RTBase/Renderer.h:30-67,
876-898: init / clear / render / getSPP / saveHDR) and talks to the GPU only through the C
ABI.  There is no CPU fallback: without the built library or without a CUDA device every
entry point raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import FlatScene  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTB200_LIB lets a developer A/B an experimental build of the same ABI (never a fallback)
LIB_PATH = os.environ.get("RTB200_LIB") or os.path.join(_HERE, "librtb200.so")

# every symbol include/rtb.h declares (tests check the built library exports all of them)
ABI_SYMBOLS = [
    "rtb_abi_version", "rtb_create", "rtb_create_multi", "rtb_device_count", "rtb_group_size", "rtb_group_info", "rtb_destroy", "rtb_last_error", "rtb_set_stream", "rtb_synchronize",
    "rtb_default_params", "rtb_set_params", "rtb_get_params", "rtb_upload_scene", "rtb_update_camera",
    "rtb_clear", "rtb_render", "rtb_render_adaptive", "rtb_render_light", "rtb_render_ir", "rtb_read_film", "rtb_write_film", "rtb_film_device_ptr", "rtb_accum_device_ptr", "rtb_set_spp", "rtb_tonemap", "rtb_get_stats",
    "rtb_film_size", "rtb_primary_hits", "rtb_trace", "rtb_visible", "rtb_shading_data", "rtb_eval_bsdf",
    "rtb_eval_light", "rtb_rng_draws",
]


class RtbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("librtb200: %s (status %d)" % (msg, code))
        self.code = code


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("%s is missing: run `python -m raytracingrenderer_b200.build` (needs nvcc); "
                              "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.rtb_abi_version.restype = i32
        L.rtb_create.argtypes = [i32, C.POINTER(vp)]
        L.rtb_create_multi.argtypes = [vp, i32, C.POINTER(vp)]
        L.rtb_device_count.restype = i32
        L.rtb_group_size.argtypes = [vp]
        L.rtb_group_info.argtypes = [vp, vp, vp, C.POINTER(u64), C.POINTER(u64)]
        L.rtb_destroy.argtypes = [vp]
        L.rtb_destroy.restype = None
        L.rtb_last_error.argtypes = [vp]
        L.rtb_last_error.restype = C.c_char_p
        L.rtb_set_stream.argtypes = [vp, vp]
        L.rtb_synchronize.argtypes = [vp]
        L.rtb_default_params.argtypes = [C.POINTER(abi.Params)]
        L.rtb_default_params.restype = None
        L.rtb_set_params.argtypes = [vp, C.POINTER(abi.Params)]
        L.rtb_get_params.argtypes = [vp, C.POINTER(abi.Params)]
        L.rtb_upload_scene.argtypes = [vp, C.POINTER(abi.SceneDesc)]
        L.rtb_update_camera.argtypes = [vp, C.POINTER(abi.Camera)]
        L.rtb_clear.argtypes = [vp]
        L.rtb_render.argtypes = [vp, u32, u32]
        L.rtb_render_adaptive.argtypes = [vp, u32, u32, u32, vp, vp]
        L.rtb_render_light.argtypes = [vp, u32, u32]
        L.rtb_render_ir.argtypes = [vp, u32, u32, u32]
        L.rtb_read_film.argtypes = [vp, vp, C.POINTER(u32)]
        L.rtb_write_film.argtypes = [vp, vp]
        L.rtb_film_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        L.rtb_accum_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        L.rtb_set_spp.argtypes = [vp, u32]
        L.rtb_tonemap.argtypes = [vp, vp, C.c_float]
        L.rtb_get_stats.argtypes = [vp, C.POINTER(abi.Stats)]
        L.rtb_film_size.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
        L.rtb_primary_hits.argtypes = [vp, i32, vp, vp, vp]
        L.rtb_trace.argtypes = [vp, i32, i32, vp, u64, vp]
        L.rtb_visible.argtypes = [vp, i32, vp, u64, vp]
        L.rtb_shading_data.argtypes = [vp, vp, vp, u64, vp]
        L.rtb_eval_bsdf.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp, vp]
        L.rtb_eval_light.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp]
        L.rtb_rng_draws.argtypes = [vp, u32, u32, u32, vp]
        if L.rtb_abi_version() != abi.ABI_VERSION:
            raise ImportError("librtb200.so ABI %d != python mirror %d" % (L.rtb_abi_version(), abi.ABI_VERSION))
        _lib = L
    return _lib


def default_params():
    p = abi.Params()
    lib().rtb_default_params(C.byref(p))
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data


class RayTracer:
    """Host mirror of the reference's ``RayTracer`` (RTBase/Renderer.h:30) on one GPU.

    ``init(scene)`` takes a FlatScene (what host/rtb_flatten.hpp produces from a live
    ``Scene``); ``render()`` is one sample per pixel like the reference's, ``render(n)``
    does n in one launch.
    """

    def __init__(self, device=0):
        """device: one CUDA device index, or a list of them / "all" for a device GROUP
        (rtb_create_multi: the film of every render call is spread over the GPUs and summed at read-out)."""
        self._L = lib()
        h = C.c_void_p()
        if device == "all" or isinstance(device, (list, tuple)):
            if device == "all":
                rc = self._L.rtb_create_multi(None, 0, C.byref(h))
            else:
                arr = (C.c_int * len(device))(*[int(d) for d in device])
                rc = self._L.rtb_create_multi(arr, len(device), C.byref(h))
            if rc != 0:
                raise RtbError(rc, self._L.rtb_last_error(None).decode())
            self._h = h
            self.device = self.group_info()["devices"][0]
        else:
            rc = self._L.rtb_create(int(device), C.byref(h))
            if rc != 0:
                raise RtbError(rc, self._L.rtb_last_error(None).decode())
            self._h = h
            self.device = int(device)
        self.scene = None
        self.width = self.height = 0

    # -- plumbing -----------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise RtbError(rc, self._L.rtb_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.rtb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def group_info(self):
        """{"devices": [...], "p2p": [...], "gathers_p2p": n, "gathers_nccl": n} of this context's device group."""
        n = self._L.rtb_group_size(self._h)
        dev, p2p = (C.c_int * n)(), (C.c_int * n)()
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self._L.rtb_group_info(self._h, dev, p2p, C.byref(a), C.byref(b)))
        return dict(devices=list(dev), p2p=list(p2p), gathers_p2p=a.value, gathers_nccl=b.value)

    def set_stream(self, cuda_stream_handle):
        self._ck(self._L.rtb_set_stream(self._h, C.c_void_p(int(cuda_stream_handle))))

    def synchronize(self):
        self._ck(self._L.rtb_synchronize(self._h))

    @property
    def params(self):
        p = abi.Params()
        self._ck(self._L.rtb_get_params(self._h, C.byref(p)))
        return p

    def set_params(self, **kw):
        p = self.params
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError("rtb_params has no field %r" % k)
            setattr(p, k, v)
        self._ck(self._L.rtb_set_params(self._h, C.byref(p)))
        return p

    # -- RayTracer surface ----------------------------------------------------------------
    def init(self, scene):
        """RayTracer::init (Renderer.h:45-63): upload the scene, allocate a cleared film."""
        d = scene.desc()
        self._ck(self._L.rtb_upload_scene(self._h, C.byref(d)))
        self.scene = scene
        self.width, self.height = scene.width, scene.height

    def update_camera(self, camera_record):
        cam = abi.Camera()
        C.memmove(C.byref(cam), np.asarray(camera_record, abi.camera_dt).tobytes(), 160)
        self._ck(self._L.rtb_update_camera(self._h, C.byref(cam)))

    def clear(self):
        self._ck(self._L.rtb_clear(self._h))

    def render(self, spp=1, spp_begin=None):
        """`spp` x RayTracer::render() (Renderer.h:876-885).  Asynchronous."""
        begin = self.getSPP() if spp_begin is None else int(spp_begin)
        self._ck(self._L.rtb_render(self._h, begin, int(spp)))

    def lightTracer(self, passes=1, pass_begin=None):
        """`passes` x RayTracer::lightTracer (Renderer.h:220-231).  Asynchronous."""
        begin = self.getSPP() if pass_begin is None else int(pass_begin)
        self._ck(self._L.rtb_render_light(self._h, begin, int(passes)))

    def instantRadiosity(self, passes=1, pass_begin=None, n_paths=50):
        """`passes` x RayTracer::instantRadiosity (Renderer.h:102-123; MAX_VPL = 50 light paths).  Asynchronous."""
        begin = self.getSPP() if pass_begin is None else int(pass_begin)
        self._ck(self._L.rtb_render_ir(self._h, begin, int(passes), int(n_paths)))

    def adaptiveRender(self, init_samples=2, min_samples=1, max_samples=10240):
        """RayTracer::adaptiveRender (Renderer.h:679-749) -> (tile_samples, tile_variance) as
        [tilesY, tilesX] arrays of the 32x32 tiles.  The film grows by one mean image, SPP by 1."""
        ty, tx = (self.height + 31) // 32, (self.width + 31) // 32
        samples = np.zeros((ty, tx), np.uint32)
        var = np.zeros((ty, tx), np.float32)
        self._ck(self._L.rtb_render_adaptive(self._h, int(init_samples), int(min_samples), int(max_samples),
                                             samples.ctypes.data, var.ctypes.data))
        return samples, var

    def write_film(self, rgb_sum):
        """Replace Film::film (running sums) — the second half of a denoise hook (Renderer.h:784-790)."""
        a = np.ascontiguousarray(rgb_sum, np.float32)
        if a.shape != (self.height, self.width, 3):
            raise ValueError("film must be [height, width, 3]")
        self._ck(self._L.rtb_write_film(self._h, a.ctypes.data))

    def getSPP(self):
        n = C.c_uint32(0)
        self._ck(self._L.rtb_read_film(self._h, None, C.byref(n)))
        return n.value

    def read_film(self, out=None):
        """Film::film as float32 [H, W, 3] running sums."""
        if out is None:
            out = np.empty((self.height, self.width, 3), "<f4")
        n = C.c_uint32(0)
        self._ck(self._L.rtb_read_film(self._h, _ptr(out), C.byref(n)))
        return out

    def film_device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._L.rtb_film_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def accum_device_ptr(self):
        """(device pointer, count) of the int64 fixed-point film sums (2^-32 units)."""
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._L.rtb_accum_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def set_spp(self, spp):
        self._ck(self._L.rtb_set_spp(self._h, int(spp)))

    def tonemap(self, exposure=1.0):
        out = np.empty((self.height, self.width, 3), np.uint8)
        self._ck(self._L.rtb_tonemap(self._h, _ptr(out), float(exposure)))
        return out

    def saveHDR(self, filename):
        """Film::save (Imaging.h:262-271): film / SPP as Radiance RGBE."""
        from .imageio import write_hdr
        spp = max(self.getSPP(), 1)
        write_hdr(filename, self.read_film() / np.float32(spp))

    def stats(self):
        s = abi.Stats()
        self._ck(self._L.rtb_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in abi.Stats._fields_}

    # -- parity entry points --------------------------------------------------------------
    def primary_hits(self, traversal=abi.TRAV_FAST, want_rays=False):
        n = self.width * self.height
        ids, t = np.empty(n, "<u4"), np.empty(n, "<f4")
        rays = np.empty(n, abi.ray_dt) if want_rays else None
        self._ck(self._L.rtb_primary_hits(self._h, traversal, _ptr(ids), _ptr(t), _ptr(rays)))
        return (ids, t, rays) if want_rays else (ids, t)

    def trace(self, rays, any_hit=False, traversal=abi.TRAV_FAST):
        rays = np.ascontiguousarray(rays, abi.ray_dt)
        hits = np.zeros(len(rays), abi.hit_dt)
        self._ck(self._L.rtb_trace(self._h, traversal, 1 if any_hit else 0, _ptr(rays), len(rays), _ptr(hits)))
        return hits

    def visible(self, p1p2, traversal=abi.TRAV_FAST):
        p = np.ascontiguousarray(p1p2, "<f4").reshape(-1, 6)
        out = np.zeros(len(p), np.uint8)
        self._ck(self._L.rtb_visible(self._h, traversal, _ptr(p), len(p), _ptr(out)))
        return out

    def shading_data(self, rays, hits):
        rays = np.ascontiguousarray(rays, abi.ray_dt)
        hits = np.ascontiguousarray(hits, abi.hit_dt)
        out = np.zeros(len(rays), abi.shading_dt)
        self._ck(self._L.rtb_shading_data(self._h, _ptr(rays), _ptr(hits), len(rays), _ptr(out)))
        return out

    def eval_bsdf(self, sd, wi, u):
        sd = np.ascontiguousarray(sd, abi.shading_dt)
        n = len(sd)
        wi = np.ascontiguousarray(wi, "<f4").reshape(n, 3)
        u = np.ascontiguousarray(u, "<f4").reshape(n, 3)
        out = dict(eval=np.zeros((n, 3), "<f4"), pdf=np.zeros(n, "<f4"), s_wi=np.zeros((n, 3), "<f4"),
                   s_f=np.zeros((n, 3), "<f4"), s_pdf=np.zeros(n, "<f4"))
        self._ck(self._L.rtb_eval_bsdf(self._h, _ptr(sd), _ptr(wi), _ptr(u), n, _ptr(out["eval"]), _ptr(out["pdf"]),
                                       _ptr(out["s_wi"]), _ptr(out["s_f"]), _ptr(out["s_pdf"])))
        return out

    def eval_light(self, light, wi, u):
        light = np.ascontiguousarray(light, "<i4")
        n = len(light)
        wi = np.ascontiguousarray(wi, "<f4").reshape(n, 3)
        u = np.ascontiguousarray(u, "<f4").reshape(n, 2)
        out = dict(p_or_wi=np.zeros((n, 3), "<f4"), emitted=np.zeros((n, 3), "<f4"), pdf=np.zeros(n, "<f4"),
                   eval=np.zeros((n, 3), "<f4"))
        self._ck(self._L.rtb_eval_light(self._h, _ptr(light), _ptr(wi), _ptr(u), n, _ptr(out["p_or_wi"]),
                                        _ptr(out["emitted"]), _ptr(out["pdf"]), _ptr(out["eval"])))
        return out

    def rng_draws(self, pixel, sample, n):
        out = np.zeros(n, "<f4")
        self._ck(self._L.rtb_rng_draws(self._h, int(pixel), int(sample), int(n), _ptr(out)))
        return out
