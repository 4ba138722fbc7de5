"""Host-side time of the pieces of one trace + detector_image step (no GPU sync inserted): what the GPU has to
wait for after the two synchronisation points of a step.  Usage on the GPU box: python tools/host_overhead.py"""
import sys, time, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np, torch
import optrace_b200 as ot
from optrace_b200 import engine, dist
import scenes
engine.ensure_init()
ot.global_options.show_warnings = False
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
N = 10_000_000
for _ in range(3):
    RT.trace(N); RT.detector_image()
torch.cuda.synchronize()
T = {}
def tick(name, t0):
    T.setdefault(name, []).append((time.perf_counter() - t0)*1e3)
for _ in range(20):
    t0 = time.perf_counter(); k = RT._geometry_state(); tick("geometry_state", t0)
    t0 = time.perf_counter(); s = RT.tracing_snapshot(); tick("tracing_snapshot", t0)
    t0 = time.perf_counter(); RT._pretrace_check(N); tick("pretrace_check", t0)
    t0 = time.perf_counter(); RT._generator_tables(); tick("generator_tables", t0)
    t0 = time.perf_counter(); RT.trace(N); tick("trace() total (incl. GPU wait)", t0)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); im = RT.detector_image(); tick("detector_image() total (incl. GPU wait)", t0)
    torch.cuda.synchronize()
for k, v in T.items():
    print(f"{k:45s} {np.median(v):8.3f} ms")
