"""Procedural flat scenes for tests that must run without the reference: a small numpy
restatement of Triangle::init (RTBase/Geometry.h:72-83), of the reference BVH builder
(Geometry.h:325-392: longest axis, centroid sort, full SAH sweep, leaves <= 2 triangles) and of
the camera set-up in loadScene (SceneLoader.h:242-260), producing raytracingrenderer_b200.abi
FlatScene objects.  Only used to MAKE inputs; results are always checked GPU vs oracle."""
import numpy as np

from raytracingrenderer_b200 import abi

F = np.float32


def _tri_records(v0, v1, v2, n0, n1, n2, uv, mat):
    n = len(v0)
    e1 = (v2 - v1).astype(F)
    e2 = (v0 - v2).astype(F)
    c = np.cross(e1, e2).astype(F)
    ln = np.sqrt((c * c).sum(axis=1, dtype=F)).astype(F)
    keep = ln > 0
    nrm = (c / np.where(ln[:, None] > 0, ln[:, None], 1)).astype(F)
    ti = np.zeros(n, abi.tri_isect_dt)
    ts = np.zeros(n, abi.tri_shade_dt)
    ti["v0"], ti["v1"], ti["v2"], ti["n"] = v0, v1, v2, nrm
    ti["d"] = (nrm * v0).sum(axis=1, dtype=F)
    ti["area"] = ln * F(0.5)
    with np.errstate(divide="ignore"):
        ti["inv_area"] = F(1.0) / (c * nrm).sum(axis=1, dtype=F)
    ti["material"] = mat
    ts["n0"], ts["n1"], ts["n2"] = n0, n1, n2
    ts["u0"], ts["u1"], ts["u2"] = uv[:, 0, 0], uv[:, 1, 0], uv[:, 2, 0]
    ts["tv0"], ts["tv1"], ts["tv2"] = uv[:, 0, 1], uv[:, 1, 1], uv[:, 2, 1]
    ts["gsign"] = np.where((n0 * nrm).sum(axis=1) > 0, 1.0, -1.0)
    return ti[keep], ts[keep]


def build_ref_bvh(ti, ts):
    """Reference-style build: returns (nodes pre-order, permuted ti, permuted ts)."""
    order = np.arange(len(ti))
    verts = np.stack([ti["v0"], ti["v1"], ti["v2"]], axis=1)  # [n,3,3]
    cent = ((verts[:, 0] + verts[:, 1] + verts[:, 2]) / F(3.0)).astype(F)
    nodes = []

    def area(mn, mx):
        s = mx - mn
        return (s[0] * s[1] + s[1] * s[2] + s[0] * s[2]) * 2.0

    def rec(lo, hi):
        idx = order[lo:hi]
        v = verts[idx].reshape(-1, 3)
        mn, mx = v.min(axis=0), v.max(axis=0)
        self_i = len(nodes)
        nodes.append(None)
        n = hi - lo
        if n <= 2:
            nodes[self_i] = (mn, ~lo, mx, n)
            return self_i
        size = mx - mn
        axis = 0
        if size[1] >= size[0] and size[1] >= size[2]:
            axis = 1
        elif size[2] >= size[0] and size[2] >= size[1]:
            axis = 2
        srt = np.argsort(cent[idx, axis], kind="stable")
        order[lo:hi] = idx[srt]
        idx = order[lo:hi]
        tmn = verts[idx].min(axis=1)
        tmx = verts[idx].max(axis=1)
        lmn, lmx = np.minimum.accumulate(tmn), np.maximum.accumulate(tmx)
        rmn, rmx = np.minimum.accumulate(tmn[::-1])[::-1], np.maximum.accumulate(tmx[::-1])[::-1]
        best, split = np.inf, 0
        for i in range(1, n):
            cost = area(lmn[i - 1], lmx[i - 1]) * i + area(rmn[i], rmx[i]) * (n - i)
            if cost < best:
                best, split = cost, i
        if split == 0:
            split = n // 2
        l = rec(lo, lo + split)
        r = rec(lo + split, hi)
        nodes[self_i] = (mn, l, mx, r)
        return self_i

    if len(ti):
        rec(0, len(ti))
    out = np.zeros(len(nodes), abi.ref_node_dt)
    for i, (mn, a, mx, b) in enumerate(nodes):
        out[i]["bmin"], out[i]["a"], out[i]["bmax"], out[i]["b"] = mn, a, mx, b
    return out, ti[order], ts[order]


def look_at_camera(frm, to, up, fov_deg, width, height):
    frm, to, up = (np.asarray(x, np.float64) for x in (frm, to, up))
    d = frm - to
    d /= np.linalg.norm(d)
    left = np.cross(up, d)
    left /= np.linalg.norm(left)
    nup = np.cross(d, left)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = left, nup, d
    view[0, 3], view[1, 3], view[2, 3] = -frm @ left, -frm @ nup, -frm @ d
    cam_to_world = np.linalg.inv(view)
    t = 1.0 / np.tan(np.radians(fov_deg) * 0.5)
    n, f = 0.001, 10000.0
    P = np.zeros((4, 4))
    P[0, 0], P[1, 1] = t / (width / height), t
    P[2, 2], P[2, 3], P[3, 2] = -f / (f - n), -(f * n) / (f - n), -1.0
    cam = np.zeros((), abi.camera_dt)
    cam["inv_proj"] = np.linalg.inv(P).astype(F).ravel()
    cam["cam_to_world"] = cam_to_world.astype(F).ravel()
    cam["origin"] = cam_to_world[:3, 3].astype(F)
    cam["width"], cam["height"] = width, height
    return cam


def assemble(ti, ts, materials, textures, texels, camera, env_tex=None, background=(0, 0, 0)):
    s = abi.FlatScene()
    s.ref_nodes, s.tri_isect, s.tri_shade = build_ref_bvh(ti, ts)
    s.materials, s.textures, s.texels = materials, textures, np.asarray(texels, F).ravel()
    s.camera = camera
    lights = []
    if env_tex is not None:
        s.background_type, s.background_tex = abi.LIGHT_ENVMAP, env_tex
        L = np.zeros((), abi.light_dt)
        L["type"], L["tex"] = abi.LIGHT_ENVMAP, env_tex
        lights.append(L)
    else:
        s.background_type, s.background_tex = abi.LIGHT_BACKGROUND, -1
        s.background_colour = np.asarray(background, F)
        if sum(background) > 0:
            L = np.zeros((), abi.light_dt)
            L["type"], L["emission"], L["tex"] = abi.LIGHT_BACKGROUND, background, -1
            lights.append(L)
    for i, t in enumerate(s.tri_isect):
        m = materials[t["material"]]
        if m["flags"] & abi.MAT_LIGHT:
            L = np.zeros((), abi.light_dt)
            L["type"], L["triangle"], L["emission"], L["area"], L["tex"] = abi.LIGHT_AREA, i, m["emission"], t["area"], -1
            lights.append(L)
    s.lights = np.array(lights, abi.light_dt) if lights else np.zeros(0, abi.light_dt)
    return s


def standard_materials():
    """One material per BSDF class (+ a layered one, an emitter and a non-two-sided rough
    dielectric), a 4x4 checker texture and 1x1 colour textures."""
    mats, texs, texels = [], [], []

    def tex(arr):
        arr = np.asarray(arr, F).reshape(-1, 3)
        t = np.zeros((), abi.texture_dt)
        t["offset"] = sum(len(x) for x in texels)
        side = int(round(np.sqrt(len(arr))))
        t["width"], t["height"] = (side, side) if side * side == len(arr) else (len(arr), 1)
        texels.append(arr)
        texs.append(t)
        return len(texs) - 1

    checker = np.where(((np.arange(16) // 4 + np.arange(16) % 4) % 2)[:, None] > 0, [0.9, 0.8, 0.2], [0.1, 0.3, 0.7])
    specs = [
        (abi.BSDF_DIFFUSE, abi.MAT_TWO_SIDED, tex(checker)),
        (abi.BSDF_MIRROR, abi.MAT_TWO_SIDED | abi.MAT_SPECULAR, tex([[0.9, 0.9, 0.9]])),
        (abi.BSDF_CONDUCTOR, abi.MAT_TWO_SIDED, tex([[0.8, 0.6, 0.3]])),
        (abi.BSDF_GLASS, abi.MAT_SPECULAR, tex([[1.0, 1.0, 1.0]])),
        (abi.BSDF_DIELECTRIC, 0, tex([[0.7, 0.7, 0.9]])),
        (abi.BSDF_ORENNAYAR, abi.MAT_TWO_SIDED, tex([[0.6, 0.2, 0.2]])),
        (abi.BSDF_PLASTIC, abi.MAT_TWO_SIDED, tex([[0.2, 0.7, 0.3]])),
        (abi.BSDF_PLASTIC, abi.MAT_TWO_SIDED | abi.MAT_LAYERED, tex([[0.5, 0.5, 0.5]])),
        (abi.BSDF_DIFFUSE, abi.MAT_TWO_SIDED | abi.MAT_LIGHT, tex([[0.0, 0.0, 0.0]])),
    ]
    for ty, fl, tx in specs:
        m = np.zeros((), abi.material_dt)
        m["type"], m["flags"], m["tex"] = ty, fl, tx
        m["int_ior"], m["ext_ior"] = 1.5, 1.0
        m["alpha"] = 0.3
        if fl & abi.MAT_LIGHT:
            m["emission"] = [9.0, 7.0, 5.0]
        mats.append(m)
    return np.array(mats, abi.material_dt), np.array(texs, abi.texture_dt), np.concatenate(texels).astype(F)


def random_scene(seed=3, n_tris=64, width=96, height=64, with_env=True):
    rng = np.random.default_rng(seed)
    mats, texs, texels0 = standard_materials()
    texels = [texels0]
    env_tex = None
    if with_env:
        env = (rng.random((8 * 16, 3)) ** 3 * 2.0).astype(F)
        t = np.zeros((), abi.texture_dt)
        t["offset"], t["width"], t["height"] = len(texels[0]), 16, 8
        texs = np.concatenate([texs, np.array([t], abi.texture_dt)])
        texels.append(env)
        env_tex = len(texs) - 1
    c = rng.uniform(-1, 1, (n_tris, 3))
    s = 0.5
    v0 = (c + rng.uniform(-s, s, (n_tris, 3))).astype(F)
    v1 = (c + rng.uniform(-s, s, (n_tris, 3))).astype(F)
    v2 = (c + rng.uniform(-s, s, (n_tris, 3))).astype(F)
    mat = rng.integers(0, len(mats) - 1, n_tris).astype(np.uint32)
    # floor (two triangles, checker diffuse) and an emitter quad above the scene
    fa, fb, fc, fd = [-2.5, -1.3, -2.5], [2.5, -1.3, -2.5], [2.5, -1.3, 2.5], [-2.5, -1.3, 2.5]
    la, lb, lc, ld = [-0.6, 1.9, -0.6], [0.6, 1.9, -0.6], [0.6, 1.9, 0.6], [-0.6, 1.9, 0.6]
    ev0 = np.array([fa, fa, la, la], F)
    ev1 = np.array([fb, fc, lb, lc], F)
    ev2 = np.array([fc, fd, lc, ld], F)
    v0, v1, v2 = np.concatenate([v0, ev0]), np.concatenate([v1, ev1]), np.concatenate([v2, ev2])
    mat = np.concatenate([mat, np.array([0, 0, len(mats) - 1, len(mats) - 1], np.uint32)])
    n = len(v0)
    gn = np.cross(v2 - v1, v0 - v2)
    gn /= np.maximum(np.linalg.norm(gn, axis=1, keepdims=True), 1e-20)
    flip = np.where(rng.random(n) < 0.3, -1.0, 1.0)[:, None]
    nn = [(gn * flip + rng.normal(scale=0.15, size=(n, 3))) for _ in range(3)]
    nn = [(x / np.linalg.norm(x, axis=1, keepdims=True)).astype(F) for x in nn]
    uv = rng.uniform(0, 3, (n, 3, 2)).astype(F)
    ti, ts = _tri_records(v0, v1, v2, nn[0], nn[1], nn[2], uv, mat)
    cam = look_at_camera([0.3, 0.4, 5.5], [0, 0, 0], [0, 1, 0], 40.0, width, height)
    return assemble(ti, ts, mats, texs, np.concatenate(texels), cam, env_tex)
