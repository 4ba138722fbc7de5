"""Throughput of every BASELINE.json config scene on one B200, next to the numpy oracle
port on the host (bounded sample).  Not the headline bench (that is bench.py on configs[1]); this table goes
into DESIGN.md / profiles.  Usage on the GPU box:  python tools/bench_configs.py [--rays N] [--cpu-rays M]"""
import argparse
import json
import sys
import time
import warnings

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import optrace_b200 as ot  # noqa: E402
from optrace_b200 import engine  # noqa: E402
from optrace_b200.scene import flatten_raytracer, detector_record  # noqa: E402
from oracle import trace_oracle as orc  # noqa: E402
import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=10_000_000)
ap.add_argument("--cpu-rays", type=int, default=100_000)
ap.add_argument("--out", default="gpurun_out/configs.json")
ap.add_argument("--only", default="", help="comma separated scene names")
args = ap.parse_args()
engine.ensure_init()
ot.global_options.show_warnings = False


def gpu_time(f, n=5, w=2):
    for _ in range(w):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0)/n


rows = []
CASES = [(n, m, False) for n, m in (("spherical_aberration", "store"), ("double_gauss", "store"), ("arizona_eye", "store"),
                                    ("image_render", "fused6"), ("cosine_surfaces", "store"), ("hurb_square", "store"),
                                    ("hurb_pinhole", "store"), ("zoo_analytic", "store"), ("zoo_numeric", "store"))]
CASES += [(n, m, True) for n, m in (("double_gauss", "store"), ("arizona_eye", "store"), ("image_render", "fused6"))]
# the reference's own published benchmark (tests/benchmark.py: 57 surfaces, 1 M rays, with and without polarisation;
# README.md:45 / docs testing.rst:94-113: 85 / 53 ms per surface per Mray on 4 cores of an i7-1360P)
if "microscope" in scenes.SCENES:
    CASES += [("microscope", "store", False), ("microscope_no_pol", "store", False)]
    scenes.SCENES["microscope_no_pol"] = lambda o: scenes.microscope(o, no_pol=True)
for name, mode, want_spec in CASES:
    if args.only and name not in args.only.split(","):
        continue
    RT = scenes.SCENES[name](ot)
    RT.use_specialised_kernels = want_spec
    spec = RT.compile() if want_spec else False
    N = args.rays
    nt = len(RT.tracing_surfaces) + 2
    if mode == "store":
        def step():
            RT.trace(N)
            return RT.detector_image()
        t = gpu_time(step)
        # trace-only time
        t_tr = gpu_time(lambda: RT.trace(N))
    else:
        RT.ITER_RAYS_STEP = N
        def step():
            return RT.iterative_render(N, pos=scenes.IMAGE_RENDER_POS)
        t = gpu_time(step, n=3, w=1)
        t_tr = None
    # host oracle on a sample of device-generated rays (same distributions)
    M = args.cpu_rays
    N_list = np.array([M//len(RT.ray_sources)]*len(RT.ray_sources))
    M = int(N_list.sum())
    rays = RT._generate(N_list, 0, M, 12345)
    f = lambda tns, k: tns.cpu().numpy().reshape((M, k), order="F") if k > 1 else tns.cpu().numpy()
    p0, s0 = f(rays.p0, 3), f(rays.s0, 3)
    pol0 = None if RT.no_pol else f(rays.pol0, 3)
    w0, wl = f(rays.w0, 1), f(rays.wl, 1)
    fs = flatten_raytracer(RT)
    hz = np.random.default_rng(0).standard_normal((fs.n_hurb, 2, M)) if fs.n_hurb else None
    t0 = time.perf_counter()
    out = orc.trace(fs, p0, s0, pol0, w0, wl, hz)
    t_cpu_trace = time.perf_counter() - t0
    rec = detector_record(RT.detectors[0].surface, "Equidistant", None)
    t0 = time.perf_counter()
    orc.detector_hits(out, rec)
    t_cpu_det = time.perf_counter() - t0
    extra = {}
    if name.startswith("microscope"):
        # the reference's metric: seconds of RT.trace(N) / len(RT.tracing_surfaces) / Mrays (tests/benchmark.py:81-86)
        extra = dict(reference_metric_ms_per_surface_per_Mray=t_tr*1e3/len(RT.tracing_surfaces)/(N/1e6),
                     published_ms_per_surface_per_Mray=dict(pol_4cores=85, no_pol_4cores=53, pol_1core=218, no_pol_1core=148,
                                                            best_no_pol_12cores=43))
    row = dict(scene=name, mode=mode, nt=nt, rays=N, no_pol=RT.no_pol, specialised=bool(spec), **extra,
               gpu_step_ms=t*1e3, gpu_trace_ms=None if t_tr is None else t_tr*1e3,
               gpu_ray_surfaces_per_s=N*(nt - 1)/t,
               gpu_trace_only_ray_surfaces_per_s=None if t_tr is None else N*(nt - 1)/t_tr,
               cpu1_ray_surfaces_per_s=M*(nt - 1)/t_cpu_trace, cpu_sample=M,
               cpu1_ms_per_surface_per_Mray=t_cpu_trace*1e3/(nt - 1)/(M/1e6))
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(args.out, "w"), indent=1)
