#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 sequential raytracing engine.

Metric (BASELINE.json): ray·surfaces/s.  Workload at every N: BASELINE.json configs[1] — the 6-element
Double Gauss (examples/double_gauss.py: 14 spherical surfaces + ring aperture + end absorber, 7 Abbe media,
5 point sources, D65, polarisation on, nt = 17 stored sections), 10 M rays PER GPU (weak scaling), store mode,
followed by the detector image (hit finding on the stored sections + XYZW binning, 4725 x 945 x 4 fp64).

One "step" = one pass of the hot path over one batch: on-device ray generation (Philox) -> trace_store kernel
-> detector_hits kernel -> render kernel (+ NCCL all-reduce of image / extent / messages when N > 1).
    value  = rays * (nt - 1) / step time, inputs (scene, sampling tables) resident in HBM, device-timed.
    e2e    = the same through the public API (Raytracer.trace + detector_image) with the scene and tables
             re-uploaded from host memory and the finished image copied back to pinned host memory each step.
    roofline: trace_store kernel, algorithmic bytes N*(nt*48 + 28) written + 68 N read, vs measured HBM peak.
    cpu_baseline / --impl reference: the numpy oracle port of the reference (oracle/trace_oracle.py) with the
             reference's own threading scheme (contiguous ray ranges per host thread) on a bounded sample.
"""
import argparse
import gc
import json
import os
import pathlib
import subprocess
import sys
import threading
import time
import warnings

ROOT = pathlib.Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

RAYS_PER_GPU = 10_000_000
CPU_SAMPLE_RAYS = 400_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU, help="rays per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--engine", default="generic", choices=["generic", "specialised"],
                    help="trace kernel build: generic (scene in the constant bank, rolled step loop) or the "
                         "scene-specialised variant of Raytracer.compile()")
    ap.add_argument("--no-compare", action="store_true", help="skip timing the other engine build")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not run the nvidia-smi clock sampler")
    ap.add_argument("--diag", "--gc-log", dest="gc_log", action="store_true",
                    help="diagnostic: per-step host times of both timed regions, phases of the slowest e2e step and garbage "
                         "collector pauses on stderr")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port with the reference's thread scheme (raytracer.py:285-286, 399-405)
# --------------------------------------------------------------------------------------------------
def cpu_threads():
    n = os.cpu_count() or 1
    if hasattr(os, "sched_getaffinity"):
        n = len(os.sched_getaffinity(0))
    return max(1, min(n, 64))      # the reference caps at 64 (misc.py:27-28)


def cpu_step(fs, det_rec, observers, bundle, n_threads):
    """trace + detector image of one bundle on the host, rays split into contiguous ranges per thread"""
    from oracle import trace_oracle as orc
    p0, s0, pol0, w0, wl = bundle
    N = p0.shape[0]
    T = max(1, min(n_threads, N//30000))
    bounds = np.linspace(0, N, T + 1).astype(int)
    outs = [None]*T

    def work(t):
        a, b = bounds[t], bounds[t + 1]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            st = orc.trace(fs, p0[a:b], s0[a:b], None if pol0 is None else pol0[a:b], w0[a:b], wl[a:b])
            outs[t] = (st, orc.detector_hits(st, det_rec))

    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    [x.start() for x in th]
    [x.join() for x in th]
    ph = np.vstack([o[1][0] for o in outs])
    w = np.concatenate([o[1][1] for o in outs])
    wlh = np.concatenate([o[1][2] for o in outs])
    ext = np.array([ph[:, 0].min(), ph[:, 0].max(), ph[:, 1].min(), ph[:, 1].max()]) if ph.shape[0] else np.zeros(4)
    e2, Nx, Ny = orc.fix_extent(ext)
    img, _ = orc.render_xyzw(observers, ph[:, 0], ph[:, 1], w, wlh, e2, Nx, Ny)
    return img, T


def cpu_bundle(N, seed=7):
    """synthetic double-gauss bundle for the host arm: 5 point sources at -50 m aimed at the lens with a 0.03
    degree isotropic cone, uniform polarisation angle, uniform wavelengths (the spectrum shape does not
    influence the cost per ray)"""
    rng = np.random.default_rng(seed)
    g = 50000.0
    deg = rng.integers(0, 5, N)*5.0
    p0 = np.zeros((N, 3))
    p0[:, 1] = -g*np.tan(np.radians(deg))
    p0[:, 2] = -g
    so = -p0/np.linalg.norm(p0, axis=1)[:, None]
    r = np.sin(np.radians(0.03))*np.sqrt(rng.random(N))
    al = rng.uniform(0, 2*np.pi, N)
    th = np.arccos(1 - r**2)
    fa = 1/np.sqrt(1 - so[:, 0]**2)
    sy = np.column_stack((np.zeros(N), -so[:, 2]*fa, so[:, 1]*fa))
    sx = np.cross(so, sy)
    s0 = np.cos(th)[:, None]*so + np.sin(th)[:, None]*(np.cos(al)[:, None]*sx + np.sin(al)[:, None]*sy)
    a = np.cross(s0, np.array([1.0, 0.0, 0.0]))
    a /= np.linalg.norm(a, axis=1)[:, None]
    b = np.cross(s0, a)
    ang = rng.uniform(0, 2*np.pi, N)
    pol0 = (a*np.cos(ang)[:, None] + b*np.sin(ang)[:, None]).astype(np.float32)
    w0 = np.full(N, 5.0/N, dtype=np.float32)
    wl = rng.uniform(380, 780, N).astype(np.float32)
    return p0, s0, pol0, w0, wl


def run_cpu(args, steps, warmup, rays):
    """fallback CPU arm: the numpy oracle port (only when the reference did not travel with the snapshot)"""
    import optrace_b200 as ot
    from optrace_b200 import color
    from optrace_b200.scene import flatten_raytracer, detector_record
    import scenes
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        RT = scenes.double_gauss(ot)
        fs = flatten_raytracer(RT)
    det = detector_record(RT.detectors[0].surface, "Equidistant", None)
    bundle = cpu_bundle(rays)
    T = cpu_threads()
    for _ in range(warmup):
        cpu_step(fs, det, color.OBSERVERS, tuple(a[:60000] if a is not None else None for a in bundle), T)
    t0 = time.perf_counter()
    used = 1
    for _ in range(steps):
        _, used = cpu_step(fs, det, color.OBSERVERS, bundle, T)
    dt = (time.perf_counter() - t0)/steps
    return rays*(fs.nt - 1)/dt, dt, used, fs.nt


REF_RAYS = 1_000_000         # N_cpu = 1e6 rays per trace, the size of tests/benchmark.py


def run_reference(steps, warmup, budget_s=150.0):
    """The UNMODIFIED reference (oracle/_ref, imported through oracle/ref_loader.py) on the host cores, timed with
    its own method (time.perf_counter around RT.trace(N), tests/benchmark.py:81-86) plus RT.detector_image():
    multithreading on, PYTHON_CPU_COUNT = min(cores, 64) (the reference rejects more, misc.py:27-28), 1e6 rays per
    step — fewer only when steps + warmup of that size would not fit the time budget on this host.
    Returns (ray*surfaces/s, s per step, threads, nt, rays per step, trace-only ray*surfaces/s)."""
    from oracle import ref_loader
    T = cpu_threads()
    os.environ["PYTHON_CPU_COUNT"] = str(T)
    ot = ref_loader.load()
    import scenes
    ot.global_options.multithreading = True
    ot.global_options.show_progress_bar = False
    ot.global_options.show_warnings = False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        RT = scenes.double_gauss(ot)
        t0 = time.perf_counter()
        RT.trace(200_000)                       # sizes the sample: rays per second of this host
        RT.detector_image()
        rate = 200_000/(time.perf_counter() - t0)
        rays = int(min(REF_RAYS, max(100_000, rate*budget_s/max(1, steps + warmup))))
        for _ in range(warmup):
            RT.trace(rays)
            RT.detector_image()
        tt = 0.0
        t0 = time.perf_counter()
        for _ in range(steps):
            a = time.perf_counter()
            RT.trace(rays)
            tt += time.perf_counter() - a
            RT.detector_image()
        dt = (time.perf_counter() - t0)/steps
    nt = RT.rays.Nt
    return rays*(nt - 1)/dt, dt, T, nt, rays, rays*(nt - 1)/(tt/steps)


def reference_or_port(steps, warmup, budget_s):
    """(value, s per step, threads, nt, rays per step, kind, sample text): the reference itself when it travelled
    with the snapshot, else the numpy oracle port (same formulas, same thread scheme)"""
    from oracle import ref_loader
    if ref_loader.available():
        v, dt, T, nt, rays, vt = run_reference(steps, warmup, budget_s)
        return v, dt, T, nt, rays, "reference", (
            f"unmodified reference (oracle/_ref), Raytracer.trace({rays}) + detector_image() per step, multithreading on, "
            f"PYTHON_CPU_COUNT={T}; {steps} steps, {dt*steps:.1f} s; trace alone {1e9/vt:.1f} ms/surface/Mray "
            f"(sections; the reference's benchmark divides by sections - 1)")
    v, dt, used, nt = run_cpu(None, steps, min(warmup, 1), CPU_SAMPLE_RAYS)
    return v, dt, used, nt, CPU_SAMPLE_RAYS, "port", (f"{CPU_SAMPLE_RAYS} rays per step, trace + detector image, numpy oracle port "
                                                     f"of the reference with its thread scheme (reference not vendored)")


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        """Started BEFORE the warm-up, so that nvidia-smi's own start-up (NVML initialisation, enumeration of every GPU
        of the box) lies outside the timed regions; mark() opens the window whose samples are reported (the timed
        regions); the polling itself is one NVML query per 100 ms.  (A/B runs with --no-clocks show no effect of the
        sampler on the timings.)"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_first(self, timeout=5.0):
        """block until the sampler has delivered its first row (its start-up is over)"""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """samples from here on count (called right before the first timed region)"""
        self.rows = self.rows[-1:]

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as td
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    import optrace_b200 as ot
    from optrace_b200 import engine, dist
    import scenes
    engine.ensure_init()
    dev = engine.device()
    N_total = args.rays*world
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        RT = scenes.double_gauss(ot)
    ot.global_options.show_warnings = False
    nt = len(RT.tracing_surfaces) + 2
    # generic kernels by default; --engine specialised selects the scene-specialised build of Raytracer.compile()
    # (in-tree cache, built by __graft_entry__.build())
    RT.use_specialised_kernels = args.engine == "specialised"
    specialised = RT.compile() if args.engine == "specialised" else False

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # ---- device-resident step: generation + trace + detector image, kernel time of the trace recorded ----
    kt, stores = [], []

    def step_resident(record):
        scene = RT._scene_handle()
        blocks = dist.shard_sources(N_list)               # this rank's slice of every source, like Raytracer.trace
        begin, end = 0, sum(c for _, _, c in blocks)
        RT._trace_count += 1
        seed = (int(RT.seed) << 20) + RT._trace_count
        rays = RT._generated(scene, N_list, blocks, 0, seed)
        e0, e1 = ev(), ev()
        # the ray store (8.2 GB) is allocated once and overwritten every step: steady-state serving pattern
        if not stores:
            stores.append(engine.DeviceStore(end - begin, scene.nt, RT.no_pol))
        store, msgs, status = engine.trace_store(scene, rays, store=stores[0], sync=False, events=(e0, e1))
        dist.allreduce_sum_(msgs)
        RT._msgs = msgs
        rays.gen_status = None
        RT.rays._attach(store, RT.ray_sources, N_list, RT.no_pol, N_total, blocks)
        RT._last_trace_snapshot = snap
        RT.check_if_rays_are_current = lambda: True
        img = RT.detector_image()
        if record:
            kt.append((e0, e1))
        return img

    from optrace_b200.ray_storage import split_rays
    N_list = dist.broadcast_ints(split_rays(N_total, [rs.power for rs in RT.ray_sources]), dev)
    snap = None
    clocks = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        clocks.start()
    # burn-in before the W warm-up steps of the contract: on a fresh box the first steps still grow the caching
    # allocator (cudaMalloc of the 8 GB ray store, image and hit buffers) and page the library in
    # (the returned image is HELD across the next step exactly like in the timed loop: otherwise the second image
    # buffer of the steady state is first allocated — a synchronous cudaMalloc of 160 MB, 5-400 ms on these boxes —
    # by the second timed step)
    for _ in range(max(0, 8 - args.warmup)):
        img = step_resident(False)
    for _ in range(args.warmup):
        img = step_resident(False)
    barrier()
    if rank == 0:
        clocks.wait_first()
        clocks.mark()
    # --diag: garbage collector passes inside the timed regions are logged (none above 0.2 ms was ever seen; the 2x
    # outliers this diagnostic was written for turned out to be cudaMalloc stalls, see the warm-up above)
    gc_pauses = []
    if args.gc_log:
        _t = [0.0]
        def _gc_cb(phase, info):
            if phase == "start":
                _t[0] = time.perf_counter()
            else:
                gc_pauses.append((info["generation"], (time.perf_counter() - _t[0])*1e3))
        gc.callbacks.append(_gc_cb)
    t0, t1 = ev(), ev()
    t0.record()
    step_wall = []
    for _ in range(args.steps):
        img = step_resident(True)
        step_wall.append(time.perf_counter())
    img._wait_device()       # N > 1: the side-stream all-reduce of the last image belongs to the timed region
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)/args.steps
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kt]))
    power = img.power()

    # for transparency: the same trace with the other engine build and with the opt-in relaxed arithmetic
    other_ms = relaxed_ms = None
    if not args.no_compare and world == 1:
        try:
            RT.use_specialised_kernels = not specialised
            RT._scene, RT._scene_key = None, None
            if RT.use_specialised_kernels:
                RT.compile()
            kt.clear()
            for k in range(4):
                step_resident(k > 0)
            torch.cuda.synchronize()
            other_ms = float(np.mean([a.elapsed_time(b) for a, b in kt]))
        except Exception as e:      # informational leg only: never lose the measurement above
            print(f"[bench] comparison engine not timed: {e}", file=sys.stderr)
        RT.use_specialised_kernels = specialised
        RT._scene, RT._scene_key = None, None
        try:
            RT.arithmetic = "relaxed"
            kt.clear()
            for k in range(4):
                step_resident(k > 0)
            torch.cuda.synchronize()
            relaxed_ms = float(np.mean([a.elapsed_time(b) for a, b in kt]))
        except Exception as e:
            print(f"[bench] relaxed arithmetic not timed: {e}", file=sys.stderr)
        RT.arithmetic = "exact"
        RT._scene, RT._scene_key = None, None
        if specialised:
            RT.compile()

    # ---- end-to-end step through the public API: uploads + trace + image + D2H of the image ----
    del RT.check_if_rays_are_current
    h2d = 0
    prev = None
    phases = []        # diagnostics (--gc-log): host time per phase [trace, materialise(prev), detector_image, download_async]

    def step_e2e():
        """one step through the public API; the image download (RenderImage.download_async, pinned staging,
        side stream) overlaps the next step's trace, the previous step's image is complete on the host before
        this function returns"""
        nonlocal h2d, prev
        RT.upload_every_trace = True          # scene record + sampling tables travel host -> device every step
        RT.deferred_status = True             # status word / message counters are collected at detector_image's own sync
        tp = [time.perf_counter()]
        RT.trace(N_total)                     # returns as soon as generator and trace kernel are queued
        tp.append(time.perf_counter())
        # bytes sent per step: kernel-parameter scene (KScene, ~30 KB), aux tables, generator tables
        h2d = 30648 + RT._scene.flat.aux.nbytes + int(RT._gen_cache[2].numel())*8
        # the previous step's image is completed on the host while this step's trace runs on the device
        out = None
        if prev is not None and rank == 0:
            out = prev._materialise()
        tp.append(time.perf_counter())
        im = RT.detector_image()
        tp.append(time.perf_counter())
        if rank == 0:       # the all-reduced image is identical on every rank: a job reads it back once
            im.download_async()
        elif prev is not None:
            prev._wait_device()               # other ranks: the reduced image is complete on the device
        tp.append(time.perf_counter())
        phases.append(np.diff(np.array(tp))*1e3)
        prev = im
        return out

    # >= 6 untimed steps, and the pipeline is NOT drained before the timed loop: the caching allocator hands the 8 GB ray
    # store, the 680 MB bundle and the image buffers out of the same large blocks, and any change of the set of live
    # tensors (a drained pipeline, the switch from the resident loop above) makes one of the next traces pay a
    # synchronous cudaMalloc of several GB (8-95 ms, seen at the third timed step in 4 of 10 runs).  The first timed
    # step therefore also completes the last warm-up image: one materialise more than steps, none less.
    for _ in range(max(6, args.warmup)):
        step_e2e()
    barrier()
    w0 = time.perf_counter()
    t0.record()
    e2e_wall = []
    phases.clear()
    for _ in range(args.steps):
        step_e2e()
        e2e_wall.append(time.perf_counter())
    # the last image is on the host (rank 0) / complete on the device (other ranks) inside the timed region as well
    out = prev._materialise() if rank == 0 else prev._wait_device()
    d2h_bytes = int(prev.transferred_bytes) if rank == 0 else 0
    t1.record()
    barrier()
    e2e_ms = max(t0.elapsed_time(t1), (time.perf_counter() - w0)*1e3)/args.steps
    if args.gc_log and rank == 0:
        d = np.diff(np.array(step_wall))*1e3
        print(f"[bench] host time between the returns of consecutive resident steps (ms): {np.round(d, 2).tolist()}", file=sys.stderr)
        d = np.diff(np.array(e2e_wall))*1e3
        print(f"[bench] host time between the returns of consecutive e2e steps (ms): {np.round(d, 2).tolist()}", file=sys.stderr)
        k = int(np.argmax(d)) + 1
        print(f"[bench] e2e phases [trace, materialise(prev), detector_image, download_async] of step {k} (ms): "
              f"{np.round(phases[k], 2).tolist()}; of step {k + 3}: {np.round(phases[min(k + 3, len(phases) - 1)], 2).tolist()}", file=sys.stderr)
        print(f"[bench] garbage collector passes (generation, ms): {[(g, round(t, 2)) for g, t in gc_pauses if t > 0.2]}", file=sys.stderr)
    clk = clocks.stop() if rank == 0 else None

    # max over ranks
    tm = torch.tensor([ms, e2e_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(tm, op=td.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms = [float(v) for v in tm.cpu()]

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        n_local = args.rays
        # the stored sections, N*(nt*48 + 28) bytes written.  The 68 B/ray read of the pre-generated
        # bundle is real traffic of this kernel too but is NOT counted as algorithmic (it exists only because
        # generation is a separate kernel); it is reported beside the roofline figure
        alg_bytes = n_local*(nt*48 + 28)
        bundle_bytes = n_local*68
        # DRAM traffic of the trace kernel from the committed `ncu --set full` capture of this same workload
        # (profiles/, dram__bytes_read.sum + dram__bytes_write.sum per launch); only valid for the default size
        traffic = None
        traffic_current = None
        profs = sorted((ROOT / "profiles").glob("r*_trace_store_final_ncu.csv"))      # newest round last
        prof = profs[-1] if profs else ROOT / "profiles" / "none"
        if prof.exists() and args.rays == RAYS_PER_GPU:
            try:
                tb = 0.0
                for ln in prof.read_text().splitlines():
                    c = ln.split(",")
                    if c[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        tb += float(c[2])*{"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[c[1]]
                    if c[0] == "engine_source_digest":
                        from optrace_b200 import build as _b
                        traffic_current = c[2] == _b.source_digest()
                traffic = tb or None
            except Exception:
                traffic = None
        achieved = alg_bytes/(kernel_ms*1e-3)/1e9
        units = N_total*(nt - 1)
        line = {
            "metric": "ray-surfaces/s", "value": units/(ms*1e-3), "unit": "ray*surface/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "double_gauss (BASELINE configs[1]): 14 spherical + ring aperture + end absorber, "
                                   "7 Abbe media, 5 point sources D65, polarisation on, store mode + detector image",
                       "rays_per_gpu": args.rays, "rays_total": N_total, "nt": nt, "sections_traced": nt - 1,
                       "image": list(img.shape), "l2": "inputs/outputs (8.9 GB per step) far larger than L2, no flush needed",
                       "parallelism": f"ray-sharded x{world}, all-reduce of image/extent/messages only",
                       "engine": ("scene-specialised trace kernel (Raytracer.compile(), cached nvcc build)" if specialised
                                  else "generic trace kernel"),
                       "trace_kernel_ms": kernel_ms,
                       ("generic_trace_kernel_ms" if specialised else "specialised_trace_kernel_ms"): other_ms,
                       "arithmetic": "exact (IEEE op-for-op, bit-identical to the reference on this scene)",
                       "relaxed_arithmetic_trace_kernel_ms": relaxed_ms, "trace_only_ray_surfaces_per_s": n_local*world*(nt - 1)/(kernel_ms*1e-3),
                       "image_power_W": power},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved/peak_gbs,
                         "traffic": traffic,
                         "traffic_source": f"profiles/{prof.name} (ncu --set full capture of this kernel on this workload; "
                                           "includes the 68 B/ray bundle read)",
                         "traffic_capture_matches_engine_sources": traffic_current,
                         "achieved_incl_bundle_read": (alg_bytes + bundle_bytes)/(kernel_ms*1e-3)/1e9,
                         "kernel": "trace_store_kernel<POL, LENS>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "e2e": {"value": units/(e2e_ms*1e-3), "unit": "ray*surface/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h_bytes,
                    "image_bytes_dense": int(out.nbytes),
                    "pipelining": "Raytracer.deferred_status: one host synchronisation per step (after the detector hit search); "
                                  "image download of step k overlaps the trace of step k+1 (depth 1); only the occupied "
                                  "32 x 32 tiles of the histogram travel (otb_tiles.cu), over NVLink and over PCIe; "
                                  "N > 1: the all-reduced image is read back on rank 0"},
            # this repo's kernels inside the two timed regions, per step: generate, trace_store, detector_hits, render
            # (+ tiles_mask, tiles_pack for the download in the e2e region; N > 1: + tiles_mask, tiles_pack, tiles_unpack
            # of the image reduce in both regions, reused by the download)
            "gpu_launches": (10 if world == 1 else 14)*args.steps,
            "clocks": clk,
        }
        if not args.no_cpu and world == 1:
            # ~15 s of host work on the same scene, the reference itself when it travelled with the snapshot
            v, dt, used, _, rays_cpu, kind, sample = reference_or_port(3, 1, 20.0)
            line["cpu_baseline"] = {"value": v, "unit": "ray*surface/s", "cores": used, "kind": kind, "sample": sample,
                                    "ms_per_surface_per_Mray": 1e9/v}
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        v, dt, used, nt, rays, kind, sample = reference_or_port(max(1, args.steps), args.warmup, 150.0)
        print(json.dumps({
            "impl": "reference", "metric": "ray-surfaces/s", "value": v, "unit": "ray*surface/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt*1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "double_gauss (BASELINE configs[1]): 14 spherical + ring aperture + end absorber, "
                                   "7 Abbe media, 5 point sources D65, polarisation on, store mode + detector image "
                                   "— on the host CPU, bounded sample per step",
                       "rays_per_step": rays, "nt": nt, "sections_traced": nt - 1},
            "cpu_baseline": {"value": v, "unit": "ray*surface/s", "cores": used, "kind": kind, "sample": sample,
                             "ms_per_surface_per_Mray": 1e9/v},
            "e2e": {"value": v, "unit": "ray*surface/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
