"""A few traces of the bench workload for profiling one engine variant under ncu.
Usage: python tools/trace_once.py [exact|relaxed] [rays]"""
import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import torch
import optrace_b200 as ot
from optrace_b200 import engine
import scenes
engine.ensure_init()
ot.global_options.show_warnings = False
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
RT.arithmetic = sys.argv[1] if len(sys.argv) > 1 else "exact"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
for _ in range(3):
    RT.trace(N)
torch.cuda.synchronize()
