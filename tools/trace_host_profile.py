"""cProfile of the host side of Raytracer.trace + detector_image in the steady state (where do the ~0.6 ms between the
kernels go).  Usage: python tools/trace_host_profile.py"""
import sys, warnings, cProfile, pstats
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import torch
import optrace_b200 as ot
import scenes
ot.global_options.show_warnings = False
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
RT.upload_every_trace = True
N = 10_000_000
prev = None
for k in range(6):
    RT.trace(N); im = RT.detector_image(); im.download_async()
    if prev is not None:
        prev._materialise()
    prev = im
pr = cProfile.Profile()
pr.enable()
for k in range(20):
    RT.trace(N); im = RT.detector_image(); im.download_async()
    prev._materialise()
    prev = im
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
