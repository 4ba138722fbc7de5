"""Trace time of fixture scenes for one engine variant (OTB_NVCC_EXTRA selects / builds the variant library).
Usage: OTB_NVCC_EXTRA="-DOTB_TRACE_THREADS_FULL=512" python tools/variant_time.py [--build-only] scene [scene ...]"""
import os, sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import optrace_b200 as ot
from optrace_b200 import userfunc
from optrace_b200.scene import flatten_raytracer
import scenes
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if "--build-only" in sys.argv:
    for name in args:
        fs = flatten_raytracer(scenes.SCENES[name](ot))
        print(name, userfunc.build_specialised_library(fs.user_funcs))
    sys.exit(0)
import torch
from optrace_b200 import engine
engine.ensure_init()
ot.global_options.show_warnings = False
N = 10_000_000
for name in args:
    RT = scenes.SCENES[name](ot)
    RT.use_specialised_kernels = False
    for _ in range(2):
        RT.trace(N)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        RT.trace(N)
    b.record()
    torch.cuda.synchronize()
    print(f"[{os.environ.get('OTB_NVCC_EXTRA', '')}] {name}: trace {a.elapsed_time(b)/4:.3f} ms", flush=True)
