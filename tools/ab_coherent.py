"""A/B timing of the bundle order (Raytracer.coherent_bundles) and of fused generation (Raytracer.fused_generation):
CUDA-event times of the generator kernel and the trace kernel on a preallocated ray store."""
import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np, torch
import optrace_b200 as ot
from optrace_b200 import engine, dist
from optrace_b200.ray_storage import split_rays
import scenes
ot.global_options.show_warnings = False
ev = lambda: torch.cuda.Event(enable_timing=True)


def run(name, N, coherent, fused, reps=6):
    RT = scenes.SCENES[name](ot)
    RT.coherent_bundles = coherent
    RT.fused_generation = fused
    RT.use_specialised_kernels = False      # generic kernels (the default build; compile() variants are opt-in)
    engine.ensure_init()
    scene = RT._scene_handle()
    N_list = split_rays(N, [rs.power for rs in RT.ray_sources])
    store = engine.DeviceStore(N, scene.nt, RT.no_pol)
    tg, tt = [], []
    for k in range(reps):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        rays = RT._generated(scene, N_list, 0, N, 1234 + k)
        e1.record()
        _, msgs, status = engine.trace_store(scene, rays, store=store, sync=False)
        e2.record()
        torch.cuda.synchronize()
        tg.append(e0.elapsed_time(e1)); tt.append(e1.elapsed_time(e2))
    w_end = store.w[N*(scene.nt - 2):N*(scene.nt - 1)]
    alive = int((w_end > 0).sum())
    print(f"{name:22s} coherent={int(coherent)} fused={int(fused)}: generate {np.mean(tg[2:]):6.3f} ms  trace {np.mean(tt[2:]):6.3f} ms"
          f"  alive at the end {alive/N:.3f}  msgs {msgs.cpu().numpy().reshape(5, -1).sum(axis=1)}", flush=True)


names = sys.argv[1:] or ["double_gauss", "arizona_eye", "spherical_aberration", "image_render", "hurb_square", "cosine_surfaces"]
for name in names:
    for coh in (False, True):
        for fused in (False, True):
            run(name, 10_000_000, coh, fused)
