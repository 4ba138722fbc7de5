"""Micro-benchmarks of the non-trace kernels (run on the GPU box): render (spread vs concentrated hits),
generate, detector hits."""
import sys, warnings, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np, torch
import optrace_b200 as ot
from optrace_b200 import engine
import scenes
lib = engine.ensure_init()
dev = engine.device()

def timeit(f, n=5, w=2):
    for _ in range(w): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n

M = 7_300_000
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.full((M,), 1e-7, dtype=torch.float32, device=dev)
wl = torch.rand(M, dtype=torch.float32, device=dev, generator=g)*400 + 380
Nx, Ny = 945, 4725
img = torch.zeros((Ny, Nx, 4), dtype=torch.float64, device=dev)
cnt = torch.zeros((Ny, Nx), dtype=torch.int32, device=dev)
for name, sx, sy in (("uniform", 1.0, 1.0), ("spot 1e-2", 1e-2, 1e-2), ("spot 1e-3", 1e-3, 1e-3), ("5 spots 3e-3", 3e-3, 3e-3)):
    x = (torch.rand(M, dtype=torch.float64, device=dev, generator=g) - 0.5)*sx
    y = (torch.rand(M, dtype=torch.float64, device=dev, generator=g) - 0.5)*sy*5
    if name.startswith("5"):
        y += (torch.arange(M, device=dev) // (M//5)).double()*0.8 - 1.6
    t = timeit(lambda: engine.render_xyzw(lib, x, y, w, wl, [-0.5, 0.5, -2.5, 2.5], Nx, Ny, img, cnt))
    print(f"render {name:14s}: {t:7.3f} ms  ({M/t/1e6:.1f} Ghit/s)")

RT = scenes.double_gauss(ot)
ot.global_options.show_warnings = False
from optrace_b200.ray_storage import split_rays
N = 10_000_000
N_list = split_rays(N, [rs.power for rs in RT.ray_sources])
RT._generate(N_list, 0, N, 1)
torch.cuda.synchronize()
t0 = time.perf_counter(); RT._generate(N_list, 0, N, 2); torch.cuda.synchronize(); print("generate 10M double_gauss:", (time.perf_counter()-t0)*1e3, "ms (wall, incl. sync)")
RT2 = scenes.arizona_eye(ot); N_list2 = np.array([N])
RT2._generate(N_list2, 0, N, 1); torch.cuda.synchronize()
t0 = time.perf_counter(); RT2._generate(N_list2, 0, N, 2); torch.cuda.synchronize(); print("generate 10M arizona (RGB image):", (time.perf_counter()-t0)*1e3, "ms")
RT.trace(N)
torch.cuda.synchronize()
t0 = time.perf_counter(); r = RT._hit_detector(); torch.cuda.synchronize(); print("hit_detector 10M:", (time.perf_counter()-t0)*1e3, "ms")
hx, hy, hw = r[0], r[1], r[2]
m = hw > 0
print("hits", int(m.sum()), "x range", float(hx[m].min()), float(hx[m].max()), "y range", float(hy[m].min()), float(hy[m].max()))
img = RT.detector_image()
c = img.counts
print("nonzero pixels", int((c > 0).sum()), "max count", int(c.max()))
