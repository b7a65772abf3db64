"""Host scene I/O: the product's stand-alone loader + reference-order builder against the
reference's own loadScene, same machine, same flattener and .rtbs writer on both sides (test tool: uses
oracle/_ref).  Writes gpurun_out/loader.json."""
import json, os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from raytracingrenderer_b200 import host_api
from oracle import ref
res = {"cpus": os.cpu_count()}
tmp = tempfile.mkdtemp()
for name in ("cornell-box", "materialball", "MaterialsScene", "coffee", "bathroom"):
    d = ref.scene_dir(name)
    host_api.load_scene(d)                       # warm the page cache for both
    t0 = time.time(); a = host_api.load_scene(d, os.path.join(tmp, "a.rtbs")); ta = time.time() - t0
    t0 = time.time(); rs = ref.RefScene(name); tl = time.time() - t0
    t0 = time.time(); b = rs.flatten(os.path.join(tmp, "b.rtbs")); tf = time.time() - t0
    same = all(getattr(a, k).tobytes() == getattr(b, k).tobytes() for k in ("ref_nodes", "tri_isect", "tri_shade", "materials", "textures", "lights", "texels"))
    res[name] = dict(triangles=int(a.n_tris), texels=int(len(a.texels)), product_load_build_flatten_s=ta, reference_load_build_s=tl,
                     reference_flatten_s=tf, byte_identical=bool(same))
    print(name, json.dumps(res[name]), flush=True)
for lg in (20, 22, 24):
    s, secs = host_api.build_soup(1 << lg, 64, 36)
    res["soup2^%d" % lg] = dict(triangles=int(s.n_tris), product_reference_order_build_s=secs)
    print("soup", lg, secs, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/loader.json", "w"), indent=1)
