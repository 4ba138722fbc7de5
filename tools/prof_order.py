import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np, torch
import optrace_b200 as ot
from optrace_b200 import engine
from optrace_b200.ray_storage import split_rays
import scenes
ot.global_options.show_warnings = False
name = sys.argv[1]
for coh in (False, True):
    RT = scenes.SCENES[name](ot)
    RT.coherent_bundles = coh
    RT.use_specialised_kernels = False
    engine.ensure_init()
    scene = RT._scene_handle()
    N = 4_000_000
    N_list = split_rays(N, [rs.power for rs in RT.ray_sources])
    store = engine.DeviceStore(N, scene.nt, RT.no_pol)
    for k in range(2):
        rays = RT._generated(scene, N_list, 0, N, 1234 + k)
        engine.trace_store(scene, rays, store=store, sync=True)
