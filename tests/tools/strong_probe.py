"""Where the time of a sample-sliced render goes on N ranks (torchrun): per-rank host timestamps and device times of
render / reduce for one scene.  usage: torchrun ... tests/tools/strong_probe.py <scene> <spp_total>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import raytracingrenderer_b200 as rtb
from raytracingrenderer_b200 import abi, host_api, distributed as D
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scene, spp = sys.argv[1], int(sys.argv[2])
s = host_api.load_scene(os.path.join("scenes", "_staged", scene))
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    rt = rtb.RayTracer(lr)
    rt.set_stream(stream.cuda_stream)
    rt.init(s)
    rt.set_params(traversal=abi.TRAV_FAST, primary_reuse=0, **D.partition_params(rank, world, "spp"))
    rt.render(4 * world, 0)
    rt.clear()
    for rep in range(2):
        rt.clear()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        rt.render(spp, 0)
        t1 = time.time()
        e[1].record()
        D.reduce_film(rt, spp)
        t2 = time.time()
        e[2].record()
        torch.cuda.synchronize()
        t3 = time.time()
        st = rt.stats()
        print("rank %d rep %d: host render call %.3f s, reduce call %.3f s, sync %.3f s | device render %.1f ms reduce %.1f ms | stats render_ms %.1f iterations %s host_syncs %s async %s"
              % (rank, rep, t1 - t0, t2 - t1, t3 - t2, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), st["render_ms"], st.get("iterations"), st.get("host_syncs"), st.get("async_renders")), flush=True)
rt.close()
dist.destroy_process_group()
