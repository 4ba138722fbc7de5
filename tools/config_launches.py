"""One trace + detector image (or fused render) of a config scene: run under
`ncu --metrics gpu__time_duration.sum` to get the per-kernel split.  Usage: python tools/config_launches.py scene [N]"""
import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import torch
import optrace_b200 as ot
from optrace_b200 import engine
import scenes
engine.ensure_init()
ot.global_options.show_warnings = False
name = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
RT = scenes.SCENES[name](ot)
RT.use_specialised_kernels = False
for _ in range(2):
    if name == "image_render":
        RT.ITER_RAYS_STEP = N
        RT.iterative_render(N, pos=scenes.IMAGE_RENDER_POS)
    else:
        RT.trace(N)
        RT.detector_image()
torch.cuda.synchronize()
