"""BASELINE.json configs at their FULL sizes on N GPUs (torchrun, one process per GPU):
  C3 arizona_eye      200 M rays, store mode + detector image (retina, Equidistant projection)
  C4 image_render     2 B rays, iterative_render, six detector positions
  C5a cosine_surfaces 50 M rays per GPU, store mode + both detectors
  C5b hurb_apertures  50 M rays per GPU (square aperture), store mode + detector image
Prints one JSON line per config (rank 0): wall time with a device synchronisation on both sides, ray.surfaces/s, image
power (must equal the power of the rays that reach the detector) and the message counters.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 \\
       tools/run_configs_full.py [--scale 1.0]"""
import argparse
import json
import os
import sys
import time
import warnings

sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0, help="scale all ray counts (smoke runs)")
args = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
import optrace_b200 as ot  # noqa: E402
from optrace_b200 import engine, dist  # noqa: E402
import scenes  # noqa: E402
engine.ensure_init()
ot.global_options.show_warnings = False
rank = dist.rank()


def sync():
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()


def report(name, N, nt, t, extra):
    if rank == 0:
        print(json.dumps(dict(config=name, n_gpus=world, rays=N, sections=nt - 1, seconds=t,
                              ray_surfaces_per_s=N*(nt - 1)/t, **extra)), flush=True)


def store_mode(name, scene, N, dets=(0,)):
    RT = scenes.SCENES[scene](ot)
    nt = len(RT.tracing_surfaces) + 2
    # two warm-up passes at full size with the timed pass's pattern of live tensors (images held while the next trace
    # runs): allocator growth, tables, image sources; a changed live set costs the next pass a synchronous cudaMalloc
    ims = None
    for _ in range(2):
        RT.trace(N)
        ims = [RT.detector_image(d) for d in dets]
        [im.power() for im in ims]
    sync()
    t0 = time.perf_counter()
    RT.trace(N)
    ims = [RT.detector_image(d) for d in dets]
    pw = [im.power() for im in ims]
    sync()
    t = time.perf_counter() - t0
    report(name, N, nt, t, dict(image_power=pw, image_shape=[list(im.shape) for im in ims],
                                 msgs=RT._msgs.sum(axis=1).tolist()))


def fused_mode(name, scene, N, pos, step):
    RT = scenes.SCENES[scene](ot)
    nt = len(RT.tracing_surfaces) + 2
    RT.ITER_RAYS_STEP = step
    [im.power() for im in RT.iterative_render(2*step, pos=pos)]         # warm-up
    sync()
    t0 = time.perf_counter()
    ims = RT.iterative_render(N, pos=pos)
    pw = [im.power() for im in ims]
    sync()
    t = time.perf_counter() - t0
    report(name, N, nt, t, dict(image_power=pw, chunks=max(1, int(N/step)), detectors=len(pos),
                                 msgs=RT._msgs.sum(axis=1).tolist()))


S = args.scale
store_mode("C3 arizona_eye, 200 M rays", "arizona_eye", int(200_000_000*S*world/8))
fused_mode("C4 image_render_many_rays, 2 B rays, 6 detector positions", "image_render", int(2_000_000_000*S*world/8),
           scenes.IMAGE_RENDER_POS, 80_000_000*world//8 if world >= 8 else 10_000_000*world)
store_mode("C5a cosine_surfaces, 50 M rays per GPU", "cosine_surfaces", int(50_000_000*S*world), dets=(0, 1))
store_mode("C5b hurb_apertures (square), 50 M rays per GPU", "hurb_square", int(50_000_000*S*world))
if world > 1:
    td.destroy_process_group()
