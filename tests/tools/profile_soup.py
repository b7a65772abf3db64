"""Tiny driver for ncu: the soup of 2^N triangles (BASELINE config 5), max_depth 0, a few spp."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import abi, host_api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
s, _ = host_api.build_soup(1 << n, 3840, 2160)
rt = rtb.RayTracer(0)
rt.init(s)
rt.set_params(max_depth=0, primary_reuse=int(os.environ.get("RTB_REUSE", "0")))
for i in range(2):
    rt.clear(); rt.render(spp, 0); rt.synchronize()
st = rt.stats()
print("soup2^%d %d spp: %.1f Msamples/s, %.1f Mrays/s, iterations %d" % (n, spp, st["samples"] / st["render_ms"] / 1e3,
      (st["closest_rays"] + st["shadow_rays"]) / st["render_ms"] / 1e3, st["iterations"]))
