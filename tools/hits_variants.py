"""Event-timed detector_hits_kernel for engine variants built with extra nvcc flags (experiment harness).
Usage: python tools/hits_variants.py --build   (CPU container)   /   python tools/hits_variants.py   (GPU box)"""
import sys, warnings, pathlib
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
from optrace_b200 import build
VARIANTS = {"base": [], "r512": ["-DOTB_DET_RESIDENT=512"], "r1024": ["-DOTB_DET_RESIDENT=1024"],
            "t256": ["-DOTB_DET_THREADS=256"], "nold": ["-DOTB_DET_PLAINLD"]}
OUT = build.ROOT / "tools" / "bin"
if "--build" in sys.argv:
    OUT.mkdir(exist_ok=True)
    build.build_library()
    base_objs = [build.CSRC / "build" / f.replace(".cu", ".o") for f in build.SOURCES if f != "otb_detect.cu"]
    for name, fl in VARIANTS.items():
        build.build_library(OUT / f"libotb_det_{name}.so", extra_flags=fl, force=True, objdir=OUT / f"obj_{name}",
                            sources=["otb_detect.cu"], extra_objects=base_objs)
        print("built", name)
    sys.exit(0)
import torch
import optrace_b200 as ot
from optrace_b200 import engine, _cabi
from optrace_b200.scene import detector_record
import scenes
ot.global_options.show_warnings = False
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
engine.ensure_init()
N = 10_000_000
RT.trace(N)
rec = detector_record(RT.detectors[0].surface, "Equidistant", None)
ev = lambda: torch.cuda.Event(enable_timing=True)
for name in VARIANTS:
    lib = _cabi.lib(OUT / f"libotb_det_{name}.so")
    for _ in range(2):
        out = engine.detector_hits(lib, RT.rays._dev, rec, 0, N)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = ev(), ev()
        a.record()
        out = engine.detector_hits(lib, RT.rays._dev, rec, 0, N)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name:8s}: {min(ts):.3f} .. {max(ts):.3f} ms   range {out[3].cpu().numpy()}")
