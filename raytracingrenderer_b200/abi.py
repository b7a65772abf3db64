"""numpy / ctypes mirrors of the POD layouts in include/rtb.h and the .rtbs container
written by host/rtb_flatten.hpp (FlatScene::save)."""
import ctypes as C

import numpy as np

ABI_VERSION = 1

camera_dt = np.dtype([
    ("inv_proj", "<f4", (16,)), ("cam_to_world", "<f4", (16,)), ("origin", "<f4", (3,)),
    ("width", "<f4"), ("height", "<f4"), ("pad_", "<f4", (3,)),
])
ref_node_dt = np.dtype([("bmin", "<f4", (3,)), ("a", "<i4"), ("bmax", "<f4", (3,)), ("b", "<i4")])
tri_isect_dt = np.dtype([
    ("v0", "<f4", (3,)), ("d", "<f4"), ("v1", "<f4", (3,)), ("inv_area", "<f4"),
    ("v2", "<f4", (3,)), ("material", "<u4"), ("n", "<f4", (3,)), ("area", "<f4"),
])
tri_shade_dt = np.dtype([
    ("n0", "<f4", (3,)), ("u0", "<f4"), ("n1", "<f4", (3,)), ("u1", "<f4"),
    ("n2", "<f4", (3,)), ("u2", "<f4"), ("tv0", "<f4"), ("tv1", "<f4"), ("tv2", "<f4"), ("gsign", "<f4"),
])
material_dt = np.dtype([
    ("type", "<u4"), ("flags", "<u4"), ("tex", "<i4"), ("int_ior", "<f4"), ("ext_ior", "<f4"),
    ("emission", "<f4", (3,)), ("alpha", "<f4"), ("eta", "<f4", (3,)), ("k", "<f4", (3,)), ("thickness", "<f4"),
])
texture_dt = np.dtype([("offset", "<u4"), ("width", "<i4"), ("height", "<i4"), ("pad_", "<i4")])
light_dt = np.dtype([
    ("type", "<u4"), ("triangle", "<u4"), ("emission", "<f4", (3,)), ("area", "<f4"), ("tex", "<i4"), ("pad_", "<i4"),
])
ray_dt = np.dtype([("o", "<f4", (3,)), ("tmax", "<f4"), ("d", "<f4", (3,)), ("pad_", "<f4")])
hit_dt = np.dtype([("id", "<u4"), ("t", "<f4"), ("alpha", "<f4"), ("beta", "<f4"), ("gamma", "<f4")])
shading_dt = np.dtype([
    ("x", "<f4", (3,)), ("wo", "<f4", (3,)), ("s_normal", "<f4", (3,)), ("g_normal", "<f4", (3,)),
    ("tu", "<f4"), ("tv", "<f4"), ("frame_u", "<f4", (3,)), ("frame_v", "<f4", (3,)), ("frame_w", "<f4", (3,)),
    ("t", "<f4"), ("material", "<i4"),
])

assert camera_dt.itemsize == 160 and ref_node_dt.itemsize == 32 and tri_isect_dt.itemsize == 64
assert tri_shade_dt.itemsize == 64 and material_dt.itemsize == 64 and texture_dt.itemsize == 16
assert light_dt.itemsize == 32 and ray_dt.itemsize == 32 and hit_dt.itemsize == 20 and shading_dt.itemsize == 100

file_header_dt = np.dtype([
    ("magic", "S8"), ("n_ref_nodes", "<u4"), ("n_tris", "<u4"), ("n_materials", "<u4"), ("n_textures", "<u4"),
    ("n_lights", "<u4"), ("background_type", "<u4"), ("background_tex", "<i4"), ("background_colour", "<f4", (3,)),
    ("pad_", "<u4", (2,)), ("n_texels", "<u8"), ("camera", camera_dt),
])
assert file_header_dt.itemsize == 224

# enums of rtb.h
BSDF_DIFFUSE, BSDF_MIRROR, BSDF_CONDUCTOR, BSDF_GLASS, BSDF_DIELECTRIC, BSDF_ORENNAYAR, BSDF_PLASTIC = range(7)
MAT_SPECULAR, MAT_TWO_SIDED, MAT_LIGHT, MAT_LAYERED = 1, 2, 4, 8
LIGHT_AREA, LIGHT_BACKGROUND, LIGHT_ENVMAP = 0, 1, 2
INT_PATH, INT_DIRECT, INT_ALBEDO, INT_NORMALS, INT_PATH_MIS = 0, 1, 2, 3, 4
RNG_MIS_BLOCK = 0x40000000
SAMPLING_STRICT, SAMPLING_IMPORTANCE = 0, 1
TRAV_EXACT, TRAV_FAST, TRAV_WIDE, TRAV_CW, TRAV_Q16 = 0, 1, 2, 3, 4
FILTER_BOX, FILTER_GAUSSIAN = 0, 1
PART_NONE, PART_SPP, PART_TILE = 0, 1, 2
SCHED_WAVEFRONT, SCHED_MEGAKERNEL = 0, 1
MISS_ID = 0xFFFFFFFF
FLT_MAX = float(np.finfo(np.float32).max)


class Camera(C.Structure):
    _fields_ = [("inv_proj", C.c_float * 16), ("cam_to_world", C.c_float * 16), ("origin", C.c_float * 3),
                ("width", C.c_float), ("height", C.c_float), ("pad_", C.c_float * 3)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("camera", Camera),
        ("ref_nodes", C.c_void_p), ("n_ref_nodes", C.c_uint32), ("n_tris", C.c_uint32),
        ("tri_isect", C.c_void_p), ("tri_shade", C.c_void_p),
        ("materials", C.c_void_p), ("n_materials", C.c_uint32), ("n_textures", C.c_uint32),
        ("textures", C.c_void_p), ("texels", C.c_void_p), ("n_texels", C.c_uint64),
        ("lights", C.c_void_p), ("n_lights", C.c_uint32),
        ("background_type", C.c_uint32), ("background_colour", C.c_float * 3), ("background_tex", C.c_int32),
    ]


class Params(C.Structure):
    _fields_ = [
        ("max_depth", C.c_int32), ("epsilon", C.c_float), ("rr_cap", C.c_float), ("integrator", C.c_int32),
        ("sampling", C.c_int32), ("traversal", C.c_int32), ("filter", C.c_int32),
        ("filter_radius", C.c_float), ("filter_alpha", C.c_float), ("seed", C.c_uint32),
        ("partition", C.c_int32), ("part_rank", C.c_int32), ("part_world", C.c_int32),
        ("cull_rel", C.c_float), ("scheduler", C.c_int32), ("primary_reuse", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("render_ms", C.c_double), ("box_tests", C.c_uint64),
        ("tri_tests", C.c_uint64), ("shadow_box_tests", C.c_uint64), ("shadow_tri_tests", C.c_uint64),
        ("extend_ms", C.c_double), ("shade_ms", C.c_double), ("shadow_ms", C.c_double),
        ("timed_iterations", C.c_uint64), ("iterations", C.c_uint64), ("host_syncs", C.c_uint64),
    ]


assert C.sizeof(Camera) == 160 and C.sizeof(Params) == 64


class FlatScene:
    """In-memory flat scene (the arrays of rtb_scene_desc) with .rtbs load/save."""

    def __init__(self):
        self.camera = np.zeros((), camera_dt)
        self.ref_nodes = np.zeros(0, ref_node_dt)
        self.tri_isect = np.zeros(0, tri_isect_dt)
        self.tri_shade = np.zeros(0, tri_shade_dt)
        self.materials = np.zeros(0, material_dt)
        self.textures = np.zeros(0, texture_dt)
        self.texels = np.zeros(0, "<f4")
        self.lights = np.zeros(0, light_dt)
        self.background_type = LIGHT_BACKGROUND
        self.background_colour = np.zeros(3, "<f4")
        self.background_tex = -1

    @property
    def width(self):
        return int(self.camera["width"])

    @property
    def height(self):
        return int(self.camera["height"])

    @property
    def n_tris(self):
        return len(self.tri_isect)

    @classmethod
    def load(cls, path):
        buf = np.fromfile(path, dtype=np.uint8)
        h = buf[:224].view(file_header_dt)[0]
        if bytes(h["magic"]) != b"RTBS0001":
            raise ValueError("%s: not an .rtbs file" % path)
        s = cls()
        s.camera = h["camera"].copy()
        s.background_type = int(h["background_type"])
        s.background_tex = int(h["background_tex"])
        s.background_colour = h["background_colour"].copy()
        off = 224

        def take(dt, n):
            nonlocal off
            nbytes = dt.itemsize * int(n)
            a = buf[off:off + nbytes].view(dt).copy()
            off += nbytes
            return a

        s.ref_nodes = take(ref_node_dt, h["n_ref_nodes"])
        s.tri_isect = take(tri_isect_dt, h["n_tris"])
        s.tri_shade = take(tri_shade_dt, h["n_tris"])
        s.materials = take(material_dt, h["n_materials"])
        s.textures = take(texture_dt, h["n_textures"])
        s.lights = take(light_dt, h["n_lights"])
        s.texels = take(np.dtype("<f4"), int(h["n_texels"]) * 3)
        if off != len(buf):
            raise ValueError("%s: trailing bytes" % path)
        return s

    def save(self, path):
        h = np.zeros((), file_header_dt)
        h["magic"] = b"RTBS0001"
        h["n_ref_nodes"], h["n_tris"] = len(self.ref_nodes), len(self.tri_isect)
        h["n_materials"], h["n_textures"], h["n_lights"] = len(self.materials), len(self.textures), len(self.lights)
        h["background_type"], h["background_tex"] = self.background_type, self.background_tex
        h["background_colour"] = self.background_colour
        h["n_texels"] = len(self.texels) // 3
        h["camera"] = self.camera
        with open(path, "wb") as f:
            f.write(h.tobytes())
            for a in (self.ref_nodes, self.tri_isect, self.tri_shade, self.materials, self.textures, self.lights,
                      self.texels):
                f.write(np.ascontiguousarray(a).tobytes())

    def desc(self):
        """ctypes rtb_scene_desc pointing into this object's arrays (keep `self` alive)."""
        d = SceneDesc()
        C.memmove(C.byref(d.camera), self.camera.tobytes(), 160)
        self._keep = [np.ascontiguousarray(a) for a in (self.ref_nodes, self.tri_isect, self.tri_shade,
                                                        self.materials, self.textures, self.texels, self.lights)]
        k = self._keep
        d.ref_nodes, d.n_ref_nodes = k[0].ctypes.data, len(k[0])
        d.n_tris, d.tri_isect, d.tri_shade = len(k[1]), k[1].ctypes.data, k[2].ctypes.data
        d.materials, d.n_materials = k[3].ctypes.data, len(k[3])
        d.n_textures, d.textures = len(k[4]), k[4].ctypes.data
        d.texels, d.n_texels = k[5].ctypes.data, len(k[5]) // 3
        d.lights, d.n_lights = k[6].ctypes.data, len(k[6])
        d.background_type = self.background_type
        for i in range(3):
            d.background_colour[i] = float(self.background_colour[i])
        d.background_tex = self.background_tex
        return d
