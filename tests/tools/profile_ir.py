"""Tiny driver for ncu: instant radiosity (k_ir_vpls + k_ir_gather) and light tracing (k_light_trace) on one scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import host_api
name = sys.argv[1] if len(sys.argv) > 1 else "cornell-box"
rt = rtb.RayTracer(0)
rt.init(host_api.load_scene(os.path.join("scenes", "_staged", name)))
for i in range(2):
    rt.clear(); rt.instantRadiosity(1, 0); rt.lightTracer(1, 0); rt.synchronize()
st = rt.stats()
print(name, "render_ms %.3f" % st["render_ms"], "rays", st["closest_rays"] + st["shadow_rays"])
