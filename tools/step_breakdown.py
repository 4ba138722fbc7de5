"""Wall-clock breakdown of one bench step (GPU box): each phase followed by a device synchronise."""
import sys, warnings, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np, torch
import optrace_b200 as ot
from optrace_b200 import engine, dist
from optrace_b200.ray_storage import split_rays
import scenes
engine.ensure_init()
ot.global_options.show_warnings = False
RT = scenes.double_gauss(ot)
RT.compile()
N = 10_000_000
N_list = split_rays(N, [rs.power for rs in RT.ray_sources])
scene = RT._scene_handle()
store = engine.DeviceStore(N, scene.nt, False)
def sync(): torch.cuda.synchronize()
T = {}
def tick(name, t0):
    sync(); T.setdefault(name, []).append((time.perf_counter()-t0)*1e3)
for it in range(6):
    sync(); t0 = time.perf_counter(); scene = RT._scene_handle(); tick("scene_handle(flatten+fingerprint)", t0)
    t0 = time.perf_counter(); rays = RT._generate(N_list, 0, N, it+1); tick("generate", t0)
    t0 = time.perf_counter(); st, msgs, status = engine.trace_store(scene, rays, store=store, sync=False); tick("trace_store", t0)
    t0 = time.perf_counter(); RT.rays._attach(st, RT.ray_sources, N_list, False, N, 0); RT._last_trace_snapshot = RT.tracing_snapshot(); tick("attach+snapshot", t0)
    t0 = time.perf_counter(); r = RT._hit_detector(); tick("_hit_detector (incl check_if_current, sync)", t0)
    hx, hy, hw, wl, extent_out = r[0], r[1], r[2], r[3], r[4]
    t0 = time.perf_counter()
    img = ot.RenderImage(extent=extent_out); img._fix_extent(); Nx, Ny = img._grid()
    data, cnt = engine.render_xyzw(scene.lib, hx, hy, hw, wl, img.extent, Nx, Ny); tick("render (alloc+zero+kernel)", t0)
for k, v in T.items():
    print(f"{k:48s} {np.mean(v[2:]):8.3f} ms")
print("sum", sum(np.mean(v[2:]) for v in T.values()))
