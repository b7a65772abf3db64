"""Render time against spp for one scene: the intercept is the per-render fixed cost (ramp-up, drain, probes).
usage: fixed_cost.py <scene> [reuse]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import abi, host_api
s = host_api.load_scene(os.path.join("scenes", "_staged", sys.argv[1]))
rt = rtb.RayTracer(0); rt.init(s)
rt.set_params(primary_reuse=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
rt.render(256, 0); rt.synchronize()
rows = []
for spp in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    best = None
    for rep in range(3):
        rt.clear(); rt.synchronize()
        t0 = time.perf_counter(); rt.render(spp, 0); rt.synchronize(); dt = time.perf_counter() - t0
        st = rt.stats()
        best = (dt, st) if best is None or dt < best[0] else best
    dt, st = best
    rows.append((spp, dt))
    print("spp %3d  %8.3f ms  iterations %4d  host_syncs %d  render_ms %.3f  extend %.3f shade %.3f shadow %.3f (sum of stage events)" % (spp, dt * 1e3, st["iterations"], st["host_syncs"], st["render_ms"], st["extend_ms"], st["shade_ms"], st["shadow_ms"]))
(b, a) = ((rows[-1][1] - rows[-2][1]) / (rows[-1][0] - rows[-2][0]), 0)
a = rows[-1][1] - b * rows[-1][0]
print("slope %.4f ms/spp, intercept %.3f ms" % (b * 1e3, a * 1e3))
