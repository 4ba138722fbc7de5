"""Fused render kernel timing (image_render scene, C4): range pass / bin pass for 1 and 6 detector positions."""
import sys, warnings, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np, torch
import optrace_b200 as ot
from optrace_b200 import engine
from optrace_b200.scene import detector_record
import scenes
engine.ensure_init(); ot.global_options.show_warnings = False
RT = scenes.image_render(ot)
if "spec" in sys.argv[1:]:
    print("specialised", RT.compile())
else:
    RT.use_specialised_kernels = False
scene = RT._scene_handle()
N = 10_000_000
rays = RT._generate(np.array([N]), 0, N, 3)
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(f, n=5, w=2):
    for _ in range(w): f()
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
for nd in (1, 2, 6):
    recs = []
    for pos in scenes.IMAGE_RENDER_POS[:nd]:
        RT.detectors[0].move_to(pos); recs.append(detector_record(RT.detectors[0].surface, "Equidistant", None))
    t_range = timeit(lambda: engine.trace_render(scene, rays, recs))
    rng = engine.trace_render(scene, rays, recs).cpu().numpy()
    imgs = [torch.zeros((945, 945, 4), dtype=torch.float64, device="cuda") for _ in range(nd)]
    cnts = [torch.zeros((945, 945), dtype=torch.int32, device="cuda") for _ in range(nd)]
    ext = [list(r) for r in rng]
    t_bin = timeit(lambda: engine.trace_render(scene, rays, recs, extents=ext, grids=[(945, 945)]*nd, imgs=imgs, cnts=cnts))
    print(f"n_det={nd}: range pass {t_range:.3f} ms, bin pass {t_bin:.3f} ms; hits/det {int(cnts[0].sum())//7}")
st = engine.DeviceStore(N, scene.nt, True)
t_store = timeit(lambda: engine.trace_store(scene, rays, store=st, sync=False))
print(f"store-mode trace of the same bundle: {t_store:.3f} ms")
RT.ITER_RAYS_STEP = N
t0 = time.perf_counter(); RT.iterative_render(4*N, pos=scenes.IMAGE_RENDER_POS); torch.cuda.synchronize(); print("iterative_render 40M rays, 6 positions:", (time.perf_counter()-t0)*1e3, "ms")
