import sys, time, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import torch, numpy as np
import optrace_b200 as ot
from optrace_b200 import engine
import scenes
engine.ensure_init()
RT = scenes.double_gauss(ot)
ot.global_options.show_warnings = False
RT.use_specialised_kernels = False
N = 10_000_000
def sync(): torch.cuda.synchronize()
prev = None
for k in range(8):
    sync(); t0 = time.perf_counter()
    RT.upload_every_trace = True
    RT.trace(N)
    t1 = time.perf_counter()
    im = RT.detector_image()
    t2 = time.perf_counter()
    im.download_async()
    t3 = time.perf_counter()
    if prev is not None: prev._materialise()
    t4 = time.perf_counter()
    prev = im
    t5 = time.perf_counter()
    sync(); t6 = time.perf_counter()
    print(f"step {k}: trace-call {1e3*(t1-t0):.2f} detimg-call {1e3*(t2-t1):.2f} dl-call {1e3*(t3-t2):.2f} wait-prev {1e3*(t4-t3):.2f} del {1e3*(t5-t4):.2f} final-sync {1e3*(t6-t5):.2f} total {1e3*(t6-t0):.2f}  pool={ {k2:len(v) for k2,v in engine._pinned_pool.items()} }")
# blocking variant
for k in range(4):
    sync(); t0 = time.perf_counter()
    RT.trace(N); im = RT.detector_image(); d = im._materialise(); sync()
    print("blocking total", 1e3*(time.perf_counter()-t0))
