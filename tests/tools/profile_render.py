"""Tiny driver for ncu: one scene, a few spp, launches k_render twice (warm + profiled)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import abi
name = sys.argv[1] if len(sys.argv) > 1 else "materialball"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
trav = {"exact": abi.TRAV_EXACT, "fast": abi.TRAV_FAST, "wide": abi.TRAV_WIDE, "cw": abi.TRAV_CW, "q16": abi.TRAV_Q16}[sys.argv[3] if len(sys.argv) > 3 else "fast"]
from raytracingrenderer_b200 import host_api
s = host_api.load_scene(os.path.join("scenes", "_staged", name))
rt = rtb.RayTracer(0)
rt.init(s)
rt.set_params(traversal=trav)
if "RTB_REUSE" in os.environ:
    rt.set_params(primary_reuse=int(os.environ["RTB_REUSE"]))
for i in range(2):
    rt.clear(); rt.render(spp, 0); rt.synchronize()
st = rt.stats()
print(name, spp, "Msamples/s %.1f  Mrays/s %.1f  box/ray %.2f tri/ray %.2f" % (st["samples"]/st["render_ms"]/1e3, (st["closest_rays"]+st["shadow_rays"])/st["render_ms"]/1e3,
      (st["box_tests"]+st["shadow_box_tests"])/(st["closest_rays"]+st["shadow_rays"]), (st["tri_tests"]+st["shadow_tri_tests"])/(st["closest_rays"]+st["shadow_rays"])))
print("  stage ms/launch: extend %.4f shade %.4f shadow %.4f  iterations %d host syncs %d launches %d" % (
    st["extend_ms"]/max(st["timed_iterations"],1), st["shade_ms"]/max(st["timed_iterations"],1), st["shadow_ms"]/max(st["timed_iterations"],1),
    st["iterations"], st["host_syncs"], st["kernel_launches"]))
