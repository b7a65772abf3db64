"""rtb_render_adaptive (RayTracer::adaptiveRender, Renderer.h:583-749) on one B200 with the reference's own
adaptiveRender() beside it (run under gpurun; test tool: uses oracle/_ref for the CPU numbers)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import host_api
from oracle import ref

res = {}
for name, with_ref in (("cornell-box_256", True), ("cornell-box", True), ("materialball", False), ("coffee", False)):
    flat = host_api.load_scene(ref.scene_dir(name))
    rt = rtb.RayTracer(0)
    rt.init(flat)
    rt.adaptiveRender(2, 1, 10240); rt.synchronize()            # warm-up (allocations)
    best = None
    for _ in range(2):
        rt.clear()
        t0 = time.time(); cnt, var = rt.adaptiveRender(2, 1, 10240); rt.synchronize(); dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    st = rt.stats()
    img = rt.read_film()
    px = rt.width * rt.height
    row = dict(res=[rt.width, rt.height], tiles=int(cnt.size), seconds=best, samples=int(st["samples"]),
               mean_samples_per_pixel=st["samples"] / px, max_tile_samples=int(cnt.max()), min_tile_samples=int(cnt.min()),
               msamples_s=st["samples"] / best / 1e6, film_mean=img.mean(axis=(0, 1)).tolist())
    rt.close()
    if with_ref:
        rs = ref.RefScene(name)
        film, rcnt, rvar, secs = rs.render_adaptive()
        rsamples = px * 2 + int((rcnt.astype(np.int64) * 1024).sum())       # padded like ours (edge tiles are full here)
        row["ref_cpu"] = dict(seconds=secs, threads=rs.hw_threads, msamples_s=rsamples / secs / 1e6,
                              film_mean=film.mean(axis=(0, 1)).tolist(), total_tile_samples=int(rcnt.sum()))
        row["speedup"] = secs / best
        row["tile_samples_total"] = int(cnt.astype(np.int64).sum())
    res[name] = row
    print(name, json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/adaptive.json", "w"), indent=1)
