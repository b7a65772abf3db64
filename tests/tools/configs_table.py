"""profiles/r01_configs_all.md from the JSON tests/tools/configs.py wrote (gpurun_out/configs.json)."""
import json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "configs.json")
d = json.load(open(src))
shutil.copyfile(src, os.path.join(ROOT, "profiles", "r01_configs_all.json"))
rows = ["# All BASELINE configs, one B200 (tests/tools/configs.py, final kernels of round 1)", "",
        "Msamples/s = W*H*spp / wall seconds of rtb_render + synchronize, scene resident.  `per-sample camera rays` traces what the",
        "reference traces (rtb_params.primary_reuse = 0, the bench.py headline); `primary-hit table` is the library default",
        "(each pixel's camera ray traced once per render call; films bit-identical, asserted by the tool).", "",
        "| config (BASELINE.json) | resolution, spp | triangles | per-sample camera rays: Msamples/s | Mrays/s | rays/sample | box tests/ray | primary-hit table: Msamples/s | reference CPU (%s threads) Msamples/s | ratio (headline) |",
        "|---|---|---|---|---|---|---|---|---|---|"]
threads = None
for k, v in d.items():
    ref = v.get("ref_cpu")
    if ref:
        threads = ref["threads"]
    rows.append("| %s | %dx%d, %d spp | %d | %.0f | %.0f | %.2f | %.1f | %.0f | %s | %s |" % (
        k, v["res"][0], v["res"][1], v["spp"], v["tris"], v["msamples_s"], v["mrays_s"], v["rays_per_sample"], v["box_per_ray"],
        v["primary_reuse"]["msamples_s"], ("%.2f" % ref["msamples_s"]) if ref else "-", ("%.0fx" % v["speedup"]) if ref else "-"))
rows[6] = rows[6] % threads
rows += ["", "soup rows: max_depth 0 (primary + one diffuse bounce); host reference-order BVH build on %s threads: %s." % (
    threads, ", ".join("%s %.1f s (upload incl. the SAH tree %.1f s)" % (k, v["host_ref_order_build_s"], v["upload_s"]) for k, v in d.items() if k.startswith("soup"))),
    "Every scene is loaded by the product host loader (librtb200_host.so: scene.json, .gem, PNG, JPEG, HDR)."]
open(os.path.join(ROOT, "profiles", "r01_configs_all.md"), "w").write("\n".join(rows) + "\n")
print("\n".join(rows))
