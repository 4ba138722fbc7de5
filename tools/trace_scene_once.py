"""A few traces of one fixture scene for profiling under ncu.  Usage: python tools/trace_scene_once.py scene [rays]"""
import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import torch
import optrace_b200 as ot
from optrace_b200 import engine
import scenes
engine.ensure_init()
ot.global_options.show_warnings = False
RT = scenes.SCENES[sys.argv[1]](ot)
RT.use_specialised_kernels = False
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
for _ in range(3):
    RT.trace(N)
torch.cuda.synchronize()
