"""GPU soak: real scenes x mode combinations, every result checked against another mode that must give
the same bits (run under gpurun).  Not a pytest module: takes a minute or two of GPU time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import abi, host_api

def render(rt, spp, **kw):
    p = rtb.default_params()
    rt.set_params(**{k: getattr(p, k) for k, _ in abi.Params._fields_})
    rt.set_params(**kw)
    rt.clear()
    rt.render(spp, 0)
    return rt.read_film().copy(), rt.stats()

bad = 0
def check(name, ok):
    global bad
    print("%-70s %s" % (name, "ok" if ok else "FAIL"), flush=True)
    bad += 0 if ok else 1

for name, spp in (("cornell-box", 96), ("materialball", 48), ("materialball_glass", 48), ("MaterialsScene", 32), ("coffee", 24), ("bathroom", 6)):
    flat = host_api.load_scene(os.path.join("scenes", "_staged", name))
    rt = rtb.RayTracer(0)
    rt.init(flat)
    base, st0 = render(rt, spp, primary_reuse=0)
    assert np.isfinite(base).all()
    a, st = render(rt, spp, primary_reuse=1)
    check("%s reuse 1 == reuse 0 (%d spp)" % (name, spp), np.array_equal(a, base) and st["samples"] == st0["samples"])
    a, _ = render(rt, spp, primary_reuse=1, traversal=abi.TRAV_EXACT) if name in ("cornell-box", "materialball") else (base, None)
    check("%s EXACT == FAST" % name, np.array_equal(a, base))
    a, _ = render(rt, spp, traversal=abi.TRAV_WIDE)
    check("%s WIDE == FAST" % name, np.array_equal(a, base))
    for trav, tn in ((abi.TRAV_CW, "CW"), (abi.TRAV_Q16, "Q16")):
        a, _ = render(rt, spp, traversal=trav)
        check("%s %s == FAST" % (name, tn), np.array_equal(a, base))
    # round 2: the same configuration again and again takes the asynchronous rtb_render path (remembered iteration
    # count); a different spp in between must not disturb it
    ok = True
    for k in range(4):
        a, st = render(rt, spp, primary_reuse=0)
        ok = ok and np.array_equal(a, base) and st["samples"] == st0["samples"]
        if k == 1:
            render(rt, max(1, spp // 3), primary_reuse=0)
    check("%s repeated renders (async path) bit-identical" % name, ok)
    # round 2: environment-selected kernels and builders give the same film
    for env in (dict(RTB_SHADOW_PERSISTENT="1"), dict(RTB_SHADOW_PERSISTENT="0", RTB_CHUNK="32"), dict(RTB_GPU_BUILD="1"),
                dict(RTB_SORT_SHADOW="1"), dict(RTB_ASYNC_RENDER="0"), dict(RTB_POOLS="1", RTB_SHADOW_ASYNC="0")):
        os.environ.update(env)
        r2 = rtb.RayTracer([0])          # a device group of one
        r2.init(flat)
        a, st = render(r2, spp, primary_reuse=0)
        r2.close()
        for k in env:
            del os.environ[k]
        check("%s %s == default" % (name, " ".join("%s=%s" % kv for kv in env.items())), np.array_equal(a, base) and st["samples"] == st0["samples"])
    # spp slices and tile slices of 3 ranks compose to the single-rank film (integer film sums)
    acc = np.zeros_like(base, dtype=np.float64)
    parts = [render(rt, spp, partition=abi.PART_TILE, part_rank=r, part_world=3)[0] for r in range(3)]
    check("%s tile partition composes" % name, np.array_equal(sum(parts), base))
    parts = [render(rt, spp, partition=abi.PART_SPP, part_rank=r, part_world=3)[0] for r in range(3)]
    check("%s spp partition composes (float sums: allclose)" % name, np.allclose(sum(parts), base, rtol=1e-5, atol=1e-5))
    # resumable: two halves == one call
    p = rtb.default_params(); rt.set_params(**{k: getattr(p, k) for k, _ in abi.Params._fields_}); rt.clear()
    rt.render(spp // 2, 0); rt.render(spp - spp // 2, spp // 2)
    check("%s two calls == one call" % name, np.array_equal(rt.read_film(), base))
    # megakernel agrees statistically-exactly (same samples, per-sample rounding differs)
    a, _ = render(rt, min(spp, 8), scheduler=abi.SCHED_MEGAKERNEL)
    b, _ = render(rt, min(spp, 8))
    check("%s megakernel ~ wavefront" % name, np.allclose(a, b, rtol=1e-4, atol=1e-4))
    for integ in (abi.INT_DIRECT, abi.INT_ALBEDO, abi.INT_NORMALS):
        a, _ = render(rt, 2, integrator=integ, primary_reuse=0)
        b, _ = render(rt, 2, integrator=integ, primary_reuse=1)
        check("%s integrator %d reuse-invariant" % (name, integ), np.array_equal(a, b) and np.isfinite(a).all())
    a, _ = render(rt, 4, integrator=abi.INT_PATH_MIS)
    check("%s MIS finite" % name, np.isfinite(a).all() and a.mean() > 0)
    p = rtb.default_params(); rt.set_params(**{k: getattr(p, k) for k, _ in abi.Params._fields_}); rt.clear()
    cnt, var = rt.adaptiveRender(2, 1, 256)
    f = rt.read_film()
    check("%s adaptive finite, counts in range" % name, np.isfinite(f).all() and cnt.min() >= 1 and cnt.max() <= 256)
    a, _ = render(rt, 4, filter=abi.FILTER_GAUSSIAN)
    check("%s gaussian filter finite" % name, np.isfinite(a).all())
    a, _ = render(rt, 4, sampling=abi.SAMPLING_IMPORTANCE)
    check("%s importance sampling finite" % name, np.isfinite(a).all())
    rt.close()
# context create / upload / render / destroy many times: device memory must come back
import torch
flat = host_api.load_scene(os.path.join("scenes", "_staged", "materialball"))
free0 = None
for i in range(12):
    rt = rtb.RayTracer(0)
    rt.init(flat)
    rt.set_params(primary_reuse=i & 1, integrator=abi.INT_PATH_MIS if i % 3 == 0 else abi.INT_PATH)
    rt.render(4, 0)
    if i % 4 == 1:
        rt.adaptiveRender(2, 1, 16)
    rt.read_film()
    rt.close()
    torch.cuda.synchronize()
    free = torch.cuda.mem_get_info()[0]
    if i == 1:
        free0 = free
check("12 x create/render/destroy: no device memory growth (%.1f MB)" % ((free0 - free) / 1e6), free0 - free < 64e6)
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
