"""Device timeline of one resident step under torchrun: CUDA events between the phases of trace + detector_image
(generator, trace kernel, message all-reduce, detector hits, all-gather + host sync, render, sparse image all-reduce),
mean over the steps, printed by rank 0.  Usage: torchrun --nproc-per-node N tools/mgpu_timeline.py"""
import os, sys, time, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np, torch, torch.distributed as td
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
import optrace_b200 as ot
from optrace_b200 import engine, dist
from optrace_b200.ray_storage import split_rays
from optrace_b200.images import RenderImage
from optrace_b200.scene import detector_record
import scenes
ot.global_options.show_warnings = False
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
engine.ensure_init()
scene = RT._scene_handle()
world = dist.world()
N_total = 10_000_000*world
N_list = split_rays(N_total, [rs.power for rs in RT.ray_sources])
blocks = dist.shard_sources(N_list)
begin, end = 0, sum(c for _, _, c in blocks)
store = engine.DeviceStore(end - begin, scene.nt, RT.no_pol)
ev = lambda: torch.cuda.Event(enable_timing=True)
names = ["generate", "trace", "msgs allreduce", "detector hits", "gather meta + host sync", "host: image setup", "render", "image reduce"]
acc = np.zeros(len(names)); host = np.zeros(2); cnt = 0
lib = scene.lib
rec = detector_record(RT.detectors[0].surface, "Equidistant", None)
for k in range(25):
    e = [ev() for _ in range(len(names) + 1)]
    t0 = time.perf_counter()
    e[0].record()
    rays = RT._generated(scene, N_list, blocks, 0, 1000 + k)
    e[1].record()
    _, msgs, status = engine.trace_store(scene, rays, store=store, sync=False)
    e[2].record()
    dist.allreduce_sum_(msgs)
    e[3].record()
    hx, hy, hw, rng, ill, st, meta = engine.detector_hits(lib, store, rec, 0, end - begin)
    e[4].record()
    r, ill_count, stt = engine.read_det_meta(meta)
    e[5].record()
    t1 = time.perf_counter()
    img = RenderImage(extent=r.copy())
    img._fix_extent()
    Nx, Ny = img._grid()
    e[6].record()
    data, cn = engine.render_xyzw(lib, hx, hy, hw, store.wl, img.extent, Nx, Ny)
    e[7].record()
    evr, tp = dist.allreduce_image_async(lib, data)
    e[8].record()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    if k >= 5:
        acc += np.array([e[i].elapsed_time(e[i + 1]) for i in range(len(names))])
        host += [t1 - t0, t2 - t1]
        cnt += 1
if dist.rank() == 0:
    print(f"world {world}: mean over {cnt} steps (each step followed by a full synchronisation)")
    for n, v in zip(names, acc/cnt):
        print(f"  {n:26s} {v:7.3f} ms")
    print(f"  sum {acc.sum()/cnt:7.3f} ms; host until the sync returned {host[0]/cnt*1e3:.3f} ms, host after it {host[1]/cnt*1e3:.3f} ms")
if world > 1:
    td.destroy_process_group()
