"""Host profile of the end-to-end step of bench.py (Raytracer.trace + detector_image + download_async through the
public API, uploads every step): cProfile over the steps, sorted by own time, plus the wall time per step.
Usage on the GPU box: python tools/e2e_profile.py [steps]"""
import cProfile, pstats, sys, time, warnings, io
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import torch
import optrace_b200 as ot
from optrace_b200 import engine
import scenes
ot.global_options.show_warnings = False
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
engine.ensure_init()
N = 10_000_000
prev = None


def step():
    global prev
    RT.upload_every_trace = True
    RT.deferred_status = True
    RT.trace(N)
    out = prev._materialise() if prev is not None else None
    im = RT.detector_image()
    im.download_async()
    prev = im
    return out


for _ in range(8):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
prev._materialise()
torch.cuda.synchronize()
print(f"plain: {(time.perf_counter() - t0)/steps*1e3:.3f} ms per e2e step")
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
for _ in range(steps):
    step()
prev._materialise()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print(f"profiled: {dt/steps*1e3:.3f} ms per e2e step")
s = io.StringIO()
ps = pstats.Stats(pr, stream=s).sort_stats("tottime")
ps.print_stats(45)
print(s.getvalue().replace("/root/repo/", ""))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(40)
print(s.getvalue().replace("/root/repo/", ""))
