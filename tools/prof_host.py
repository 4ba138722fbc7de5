import sys, warnings, cProfile, pstats
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
N = 10_000_000
for _ in range(3):
    RT.trace(N); RT.detector_image()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(40):
    RT.trace(N); RT.detector_image()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
