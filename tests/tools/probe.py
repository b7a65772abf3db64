"""Dev probe (run under gpurun): GPU vs oracle/_ref on the staged scenes."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import abi
from oracle import ref

scenes = sys.argv[1:] or ["cornell-box", "MaterialsScene", "materialball", "coffee", "bathroom"]
out = {}
for name in scenes:
    t0 = time.time()
    rs = ref.RefScene(name)
    flat = rs.flatten("/tmp/%s.rtbs" % name)
    t1 = time.time()
    rt = rtb.RayTracer(0)
    rt.init(flat)
    t2 = time.time()
    rid, rtt = rs.primary_hits()
    t3 = time.time()
    res = {"load_s": t1 - t0, "upload_s": t2 - t1, "ref_primary_s": t3 - t2}
    for tname, trav in (("exact", abi.TRAV_EXACT), ("fast", abi.TRAV_FAST)):
        ta = time.time()
        ids, tt = rt.primary_hits(trav)
        tb = time.time()
        res[tname] = {"id_mismatch": int((ids != rid).sum()), "t_mismatch": int((tt.view(np.uint32) != rtt.view(np.uint32)).sum()), "s": tb - ta}
    spp = 16
    for tname, trav in (("fast", abi.TRAV_FAST), ("exact", abi.TRAV_EXACT)):
        rt.set_params(traversal=trav)
        rt.clear()
        rt.render(2); rt.synchronize(); rt.clear()
        ta = time.time(); rt.render(spp); rt.synchronize(); tb = time.time()
        film = rt.read_film() / spp
        st = rt.stats()
        res["render_" + tname] = {"s": tb - ta, "msamples_s": st["samples"] / (tb - ta) / 1e6,
            "mrays_s": (st["closest_rays"] + st["shadow_rays"]) / (tb - ta) / 1e6, "mean": film.mean(axis=(0, 1)).tolist(),
            "rays_per_sample": (st["closest_rays"] + st["shadow_rays"]) / max(st["samples"], 1),
            "box_per_ray": (st["box_tests"] + st["shadow_box_tests"]) / max(st["closest_rays"] + st["shadow_rays"], 1),
            "tri_per_ray": (st["tri_tests"] + st["shadow_tri_tests"]) / max(st["closest_rays"] + st["shadow_rays"], 1),
            "render_ms": st["render_ms"], "nan": int(np.isnan(film).sum())}
    rspp = 4 if name != "bathroom" else 1
    rf, n, secs = rs.render(rspp, 0)
    res["ref"] = {"spp": n, "s": secs, "msamples_s": rs.width * rs.height * n / secs / 1e6, "mean": (rf / n).mean(axis=(0, 1)).tolist(), "threads": rs.hw_threads}
    out[name] = res
    print(name, json.dumps(res), flush=True)
    rt.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
