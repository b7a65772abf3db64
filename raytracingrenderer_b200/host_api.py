"""Python front end of librtb200_host.so — the stand-alone host layer (host/rtb_scene.hpp):
`load_scene(dir)` is the reference's `loadScene` (RTBase/SceneLoader.h:237) followed by the
flattener, `build_soup` the synthetic random-triangle scene of SURVEY 8d cfg 5.  No reference
code is involved; tests/test_host_cpu.py checks the result byte for byte against the reference's
own loader."""
import ctypes as C
import os
import tempfile

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "librtb200_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(HOST_LIB_PATH):
            raise ImportError("%s is missing: run `python -m raytracingrenderer_b200.build`" % HOST_LIB_PATH)
        L = C.CDLL(HOST_LIB_PATH)
        L.rtbh_last_error.restype = C.c_char_p
        L.rtbh_load_scene_rtbs.argtypes = [C.c_char_p, C.c_char_p]
        L.rtbh_matrix_invert.argtypes = [C.c_void_p, C.c_void_p]
        L.rtbh_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.rtbh_build_soup.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_double)]
        L.rtbh_sort_selftest.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.rtbh_decode_hdr.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_uint64]
        L.rtbh_write_hdr.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.rtbh_write_png.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.rtbh_decode_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


def load_scene(scene_dir, rtbs_path=None):
    """loadScene(scene_dir) -> FlatScene (also written to rtbs_path if given)."""
    tmp = None
    if rtbs_path is None:
        fd, tmp = tempfile.mkstemp(suffix=".rtbs")
        os.close(fd)
        rtbs_path = tmp
    try:
        rc = lib().rtbh_load_scene_rtbs(scene_dir.encode(), rtbs_path.encode())
        if rc != 0:
            raise RuntimeError("loadScene(%s): %s" % (scene_dir, lib().rtbh_last_error().decode()))
        return abi.FlatScene.load(rtbs_path)
    finally:
        if tmp and os.path.exists(tmp):
            os.unlink(tmp)


def matrix_invert(m):
    a = np.ascontiguousarray(m, "<f4").reshape(16)
    out = np.zeros(16, "<f4")
    lib().rtbh_matrix_invert(a.ctypes.data, out.ctypes.data)
    return out


def camera(frm, to, up, fov, width, height):
    f, t, u = (np.ascontiguousarray(x, "<f4") for x in (frm, to, up))
    out = np.zeros(35, "<f4")
    lib().rtbh_camera(f.ctypes.data, t.ctypes.data, u.ctypes.data, float(fov), int(width), int(height), out.ctypes.data)
    cam = np.zeros((), abi.camera_dt)
    cam["inv_proj"], cam["cam_to_world"], cam["origin"] = out[:16], out[16:32], out[32:35]
    cam["width"], cam["height"] = width, height
    return cam


def build_soup(n_tris, width=3840, height=2160, rtbs_path=None):
    """-> (FlatScene, BVH build seconds)"""
    tmp = None
    if rtbs_path is None:
        fd, tmp = tempfile.mkstemp(suffix=".rtbs")
        os.close(fd)
        rtbs_path = tmp
    secs = C.c_double(0)
    try:
        rc = lib().rtbh_build_soup(int(n_tris), int(width), int(height), rtbs_path.encode(), C.byref(secs))
        if rc != 0:
            raise RuntimeError("build_soup: %s" % lib().rtbh_last_error().decode())
        return abi.FlatScene.load(rtbs_path), secs.value
    finally:
        if tmp and os.path.exists(tmp):
            os.unlink(tmp)


def sort_selftest(keys, par=3):
    """True if the builder's parallel sort gives std::sort's permutation on `keys` (float32)."""
    k = np.ascontiguousarray(keys, np.float32)
    return bool(lib().rtbh_sort_selftest(k.ctypes.data, len(k), int(par)))


def decode_image(path):
    """PNG / JPEG -> uint8 [H, W, channels] with the loader's own decoders; raises on failure."""
    w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
    if lib().rtbh_decode_image(path.encode(), C.byref(w), C.byref(h), C.byref(c), None, 0) != 0:
        raise RuntimeError("cannot decode " + path)
    out = np.zeros((h.value, w.value, c.value), np.uint8)
    if lib().rtbh_decode_image(path.encode(), C.byref(w), C.byref(h), C.byref(c), out.ctypes.data, out.size) != 0:
        raise RuntimeError("cannot decode " + path)
    return out


def decode_hdr(path):
    """Radiance .hdr -> float32 [H, W, 3] with the loader's own decoder; raises on failure."""
    w, h = C.c_int(0), C.c_int(0)
    if lib().rtbh_decode_hdr(path.encode(), C.byref(w), C.byref(h), None, 0) != 0:
        raise RuntimeError("cannot decode " + path)
    out = np.zeros((h.value, w.value, 3), np.float32)
    lib().rtbh_decode_hdr(path.encode(), C.byref(w), C.byref(h), out.ctypes.data, out.size)
    return out


def write_hdr(path, rgb):
    """float32 [H, W, 3] -> Radiance .hdr with the stand-alone program's writer (Film::save)."""
    a = np.ascontiguousarray(rgb, np.float32)
    if lib().rtbh_write_hdr(path.encode(), a.shape[1], a.shape[0], a.ctypes.data) != 0:
        raise IOError("cannot write " + path)


def write_png(path, img):
    """uint8 [H, W, C] (C = 1, 3 or 4) -> PNG with the stand-alone program's writer (savePNG)."""
    a = np.ascontiguousarray(img, np.uint8)
    if lib().rtbh_write_png(path.encode(), a.shape[1], a.shape[0], a.shape[2], a.ctypes.data) != 0:
        raise IOError("cannot write " + path)
