"""Event-timed trace kernel of the bench workload for engine variants built with extra nvcc flags (experiment harness).
Usage: python tools/lens_variants.py --build (CPU container) / python tools/lens_variants.py NAME (GPU box, one variant per process)"""
import sys, warnings, pathlib
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
from optrace_b200 import build
VARIANTS = {"base": [], "mb5": ["-DOTB_LENS_MINBLOCKS=5"], "t64b9": ["-DOTB_TRACE_THREADS=64", "-DOTB_LENS_MINBLOCKS=9"],
            "t64b10": ["-DOTB_TRACE_THREADS=64", "-DOTB_LENS_MINBLOCKS=10"], "t96b6": ["-DOTB_TRACE_THREADS=96", "-DOTB_LENS_MINBLOCKS=6"],
            "t256b2": ["-DOTB_TRACE_THREADS=256", "-DOTB_LENS_MINBLOCKS=2"]}
OUT = build.ROOT / "tools" / "bin"
if "--build" in sys.argv:
    import concurrent.futures
    OUT.mkdir(exist_ok=True)
    build.build_library()
    base_objs = [build.CSRC / "build" / f.replace(".cu", ".o") for f in build.SOURCES if f != "otb_trace.cu"]
    def one(item):
        name, fl = item
        build.build_library(OUT / f"libotb_lens_{name}.so", extra_flags=fl, force=True, objdir=OUT / f"obj_l_{name}",
                            sources=["otb_trace.cu"], extra_objects=base_objs)
        return name
    with concurrent.futures.ThreadPoolExecutor(4) as ex:
        for n in ex.map(one, VARIANTS.items()):
            print("built", n)
    sys.exit(0)
name = sys.argv[1]
from optrace_b200 import _cabi
_cabi.LIB_PATH = OUT / f"libotb_lens_{name}.so"
import torch
import optrace_b200 as ot
from optrace_b200 import engine, dist
from optrace_b200.ray_storage import split_rays
import scenes
ot.global_options.show_warnings = False
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
engine.ensure_init()
N = 10_000_000
scene = RT._scene_handle()
N_list = split_rays(N, [rs.power for rs in RT.ray_sources])
store = engine.DeviceStore(N, scene.nt, RT.no_pol)
ev = lambda: torch.cuda.Event(enable_timing=True)
ts = []
for k in range(8):
    rays = RT._generated(scene, N_list, dist.shard_sources(N_list), 0, 100 + k)
    a, b = ev(), ev()
    engine.trace_store(scene, rays, store=store, sync=False, events=(a, b))
    torch.cuda.synchronize()
    if k >= 3:
        ts.append(a.elapsed_time(b))
print(f"{name}: trace kernel {min(ts):.3f} .. {max(ts):.3f} ms")
