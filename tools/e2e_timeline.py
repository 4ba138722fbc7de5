"""Host-side timeline of the end-to-end step of bench.py (Raytracer.trace + detector_image + pipelined download):
wall time of every phase per step, to see where the host waits.  Usage: python tools/e2e_timeline.py [steps]"""
import sys, time, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np, torch
import optrace_b200 as ot
from optrace_b200 import engine
import scenes
ot.global_options.show_warnings = False
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
RT.upload_every_trace = True
N = 10_000_000
prev = None
rows = []
for k in range(steps + 5):
    t0 = time.perf_counter()
    RT.trace(N)
    t1 = time.perf_counter()
    im = RT.detector_image()
    t2 = time.perf_counter()
    im.download_async()
    t3 = time.perf_counter()
    if prev is not None:
        prev._materialise()
    t4 = time.perf_counter()
    prev = im
    rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
torch.cuda.synchronize()
r = np.array(rows[5:])*1e3
print("phase            mean    median   max   [ms]")
for i, nm in enumerate(("trace()", "detector_image()", "download_async()", "materialise(prev)", "step total")):
    print(f"{nm:18s} {r[:, i].mean():7.3f} {np.median(r[:, i]):7.3f} {r[:, i].max():7.3f}")
print("per-step totals:", np.round(r[:, 4], 2))
