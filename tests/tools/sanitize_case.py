"""Small end-to-end case for compute-sanitizer (one tool per gpurun call): the procedural test
scene through every traversal, integrator, partition and the parity entry points."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import abi
import refbvh, raysets
s = refbvh.random_scene()
rt = rtb.RayTracer(0)
rt.init(s)
for trav in (abi.TRAV_EXACT, abi.TRAV_FAST, abi.TRAV_WIDE, abi.TRAV_CW, abi.TRAV_Q16):
    ids, t, rays = rt.primary_hits(trav, want_rays=True)
    hits = rt.trace(rays, traversal=trav)
    sets = raysets.mixed_set(rays, hits, 1, 500)
    rt.trace(sets["closest"], traversal=trav); rt.trace(sets["anyhit"], any_hit=True, traversal=trav); rt.visible(sets["segments"], traversal=trav)
    for integ in (abi.INT_PATH, abi.INT_DIRECT, abi.INT_ALBEDO, abi.INT_NORMALS):
        for sched in (abi.SCHED_WAVEFRONT, abi.SCHED_MEGAKERNEL):
            rt.set_params(traversal=trav, integrator=integ, scheduler=sched, partition=abi.PART_NONE, part_world=1, part_rank=0)
            rt.clear(); rt.render(3, 0); f = rt.read_film()
            assert np.isfinite(f).all()
rt.set_params(traversal=abi.TRAV_FAST, integrator=abi.INT_PATH, scheduler=abi.SCHED_WAVEFRONT, sampling=abi.SAMPLING_IMPORTANCE)
for part in (abi.PART_SPP, abi.PART_TILE):
    rt.set_params(partition=part, part_rank=1, part_world=3); rt.clear(); rt.render(5, 0); rt.read_film()
rt.set_params(filter=abi.FILTER_GAUSSIAN); rt.read_film(); rt.tonemap()
sd = rt.shading_data(sets["closest"], rt.trace(sets["closest"]))
m = sd["material"] >= 0
n = int(m.sum())
rng = np.random.default_rng(0)
rt.eval_bsdf(sd[m], raysets.unit(rng.normal(size=(n, 3))), rng.random((n, 3), dtype=np.float32))
rt.eval_light(rng.integers(0, len(s.lights), n).astype(np.int32), raysets.unit(rng.normal(size=(n, 3))), rng.random((n, 2), dtype=np.float32))
n_launch = rt.stats()["kernel_launches"]
rt.close()
# round 2: ray binning, the persistent any-hit kernel, small chunks, the device-built tree, MIS + importance sampling,
# adaptive / light / IR on a device group of one, film write-back
for env in (dict(RTB_SORT_SHADOW="1", RTB_SORT_EXTEND="1"), dict(RTB_SHADOW_PERSISTENT="1", RTB_CHUNK="32"), dict(RTB_GPU_BUILD="1"),
            dict(RTB_CW_STAGE_KB="0", RTB_CW_SHADOW="0")):
    os.environ.update(env)
    r2 = rtb.RayTracer([0])
    r2.init(s)
    for trav in (abi.TRAV_FAST, abi.TRAV_CW, abi.TRAV_Q16):
        for integ in (abi.INT_PATH, abi.INT_PATH_MIS):
            r2.set_params(traversal=trav, integrator=integ, sampling=abi.SAMPLING_IMPORTANCE)
            r2.clear(); r2.render(3, 0)
            assert np.isfinite(r2.read_film()).all()
    r2.set_params(traversal=abi.TRAV_FAST, integrator=abi.INT_PATH)
    r2.clear(); r2.adaptiveRender(2, 1, 16); r2.lightTracer(2); r2.instantRadiosity(1); f = r2.read_film()
    r2.write_film(f); r2.tonemap()
    n_launch += r2.stats()["kernel_launches"]
    r2.close()
    for k in env:
        del os.environ[k]
print("sanitize case ok", n_launch, "launches")
