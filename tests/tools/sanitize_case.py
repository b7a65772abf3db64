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
print("sanitize case ok", rt.stats()["kernel_launches"], "launches")
rt.close()
