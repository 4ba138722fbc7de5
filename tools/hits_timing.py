"""detector_hits_kernel in the pipeline (right after the trace kernel) vs repeated on the same store.
Usage on the GPU box: python tools/hits_timing.py"""
import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import torch
import optrace_b200 as ot
from optrace_b200 import engine
from optrace_b200.scene import detector_record
import scenes
ot.global_options.show_warnings = False
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
engine.ensure_init()
N = 10_000_000
ev = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(3):
    RT.trace(N); RT.detector_image()
rec = detector_record(RT.detectors[0].surface, "Equidistant", None)
lib = RT._scene.lib
for rep in range(3):
    RT.trace(N)
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record()
    out = engine.detector_hits(lib, RT.rays._dev, rec, 0, N)
    b.record()
    torch.cuda.synchronize()
    c.record()
    out = engine.detector_hits(lib, RT.rays._dev, rec, 0, N)
    d.record()
    torch.cuda.synchronize()
    print(f"hits after a completed trace: {a.elapsed_time(b):.3f} ms, repeated: {c.elapsed_time(d):.3f} ms, status {int(RT.rays._dev.status.item())}")
# back to back with the trace kernel still running (no host sync in between)
scene = RT._scene_handle()
import numpy as np
from optrace_b200 import dist
from optrace_b200.ray_storage import split_rays
N_list = split_rays(N, [rs.power for rs in RT.ray_sources])
for rep in range(3):
    rays = RT._generated(scene, N_list, dist.shard_sources(N_list), 0, 77 + rep)
    store, msgs, status = engine.trace_store(scene, rays, sync=False)
    a, b = ev(), ev()
    a.record()
    out = engine.detector_hits(lib, store, rec, 0, N)
    b.record()
    torch.cuda.synchronize()
    print(f"hits queued behind the trace kernel: {a.elapsed_time(b):.3f} ms")
