"""GPU timeline of the end-to-end step (kernels and copies as CUPTI sees them, via torch.profiler): the idle gaps between
consecutive GPU activities and what surrounds them.  Usage on the GPU box: python tools/gpu_gaps.py [e2e|resident]"""
import sys, warnings, json, tempfile, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import torch
from torch.profiler import profile, ProfilerActivity
import optrace_b200 as ot
from optrace_b200 import engine
import scenes
ot.global_options.show_warnings = False
mode = sys.argv[1] if len(sys.argv) > 1 else "e2e"
RT = scenes.double_gauss(ot)
RT.use_specialised_kernels = False
engine.ensure_init()
N = 10_000_000
prev = None


def step():
    global prev
    RT.upload_every_trace = mode == "e2e"
    RT.deferred_status = mode == "e2e"
    RT.trace(N)
    if mode == "e2e" and prev is not None:
        prev._materialise()
    im = RT.detector_image()
    if mode == "e2e":
        im.download_async()
    prev = im


for _ in range(8):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "otb_trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
gpu = sorted((e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e), key=lambda e: e["ts"])
t_end = None
print(f"{'start ms':>9} {'dur us':>8} {'gap us':>8}  stream  name")
t0 = gpu[0]["ts"]
busy = 0.0
for e in gpu:
    gap = (e["ts"] - t_end) if t_end is not None else 0.0
    name = e["name"][:60]
    if e["dur"] > 20 or gap > 20:
        print(f"{(e['ts'] - t0)/1e3:9.3f} {e['dur']:8.1f} {gap:8.1f}  {e['args'].get('stream', '?'):>6}  {name}")
    t_end = max(t_end or 0, e["ts"] + e["dur"])
    busy += e["dur"]
print(f"span {(t_end - t0)/1e3:.3f} ms for 4 steps, sum of activity durations {busy/1e3:.3f} ms")
