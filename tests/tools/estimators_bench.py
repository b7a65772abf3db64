"""The estimators the reference ships switched off, on one B200 with the reference's own timings beside them
(run under gpurun; test tool: uses oracle/_ref for the CPU numbers).  Writes gpurun_out/estimators.json."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import abi, host_api
from oracle import ref

def timed(fn, rt, reps=2):
    best = None
    for _ in range(reps):
        rt.clear(); t0 = time.time(); fn(); rt.synchronize(); dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    return best, rt.stats()

res = {}
for name in ("cornell-box_256", "cornell-box", "coffee"):
    flat = host_api.load_scene(ref.scene_dir(name))
    rt = rtb.RayTracer(0); rt.init(flat)
    px = rt.width * rt.height
    row = {"res": [rt.width, rt.height]}
    rt.render(4, 0); rt.lightTracer(1); rt.instantRadiosity(1); rt.synchronize()
    for label, integ in (("path", abi.INT_PATH), ("path_mis", abi.INT_PATH_MIS)):
        rt.set_params(integrator=integ)
        dt, st = timed(lambda: rt.render(64, 0), rt)
        row[label] = {"spp": 64, "seconds": dt, "msamples_s": st["samples"] / dt / 1e6}
    rt.set_params(integrator=abi.INT_PATH)
    dt, st = timed(lambda: rt.lightTracer(16, 0), rt)
    row["light_tracing"] = {"passes": 16, "seconds_per_pass": dt / 16, "mpaths_s": st["samples"] / dt / 1e6}
    dt, st = timed(lambda: rt.instantRadiosity(8, 0), rt)
    row["instant_radiosity"] = {"passes": 8, "seconds_per_pass": dt / 8, "shadow_mrays_per_pass": st["shadow_rays"] / 8e6,
                                "mrays_s": (st["closest_rays"] + st["shadow_rays"]) / dt / 1e6}
    rt.close()
    if name == "cornell-box_256":
        rs = ref.RefScene(name)
        rs.render(1, 0, fresh=True)
        _, _, secs = rs.render(8, 0, fresh=True)
        row["path"]["ref_cpu_msamples_s"] = px * 8 / secs / 1e6
        rm = ref.RefScene(name, "_mis")
        rm.render(1, 0, fresh=True)
        _, _, secs = rm.render(8, 0, fresh=True)
        row["path_mis"]["ref_cpu_msamples_s"] = px * 8 / secs / 1e6
        _, secs = rs.render_light(8)
        row["light_tracing"]["ref_cpu_seconds_per_pass"] = secs / 8
        row["light_tracing"]["ref_cpu_mpaths_s"] = px * 8 / secs / 1e6
        _, secs, nv = rs.render_ir(4)
        row["instant_radiosity"]["ref_cpu_seconds_per_pass"] = secs / 4
        row["ref_threads"] = rs.hw_threads
    res[name] = row
    print(name, json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/estimators.json", "w"), indent=1)
