"""Worker of tests/test_gpu_multi.py (one process per GPU under torchrun): the sharded, all-reduced results of N
ranks against the single-GPU result on the SAME injected bundle."""
import json
import os
import pathlib
import sys
import warnings

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
warnings.simplefilter("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))

import optrace_b200 as ot  # noqa: E402
from optrace_b200 import dist  # noqa: E402
import golden_util as gu  # noqa: E402
import scenes  # noqa: E402

ot.global_options.show_warnings = False
rank, world = dist.rank(), dist.world()
report = {}


def fixture(name):
    big = ROOT / "tests" / "golden_large" / f"{name}.npz"
    return dict(np.load(big)) if big.exists() else gu.load(name)


for name in ("double_gauss", "spherical_aberration", "hurb_square", "arizona_eye"):
    g = fixture(name)
    p0, s0, pol0, w0, wl, hz = gu.bundle(g)
    N = p0.shape[0]
    # reference: this process alone (no sharding, no collectives) on the full bundle
    with dist.local_mode():
        RT1 = scenes.SCENES[name](ot)
        RT1.trace_rays(p0, s0, pol0, w0, wl, hurb_z=hz, N_list=g["N_list"])
        im1 = RT1.detector_image()
        cnt1, dat1, ext1, msgs1 = im1.counts.copy(), im1.data.copy(), np.array(im1.extent), RT1._msgs.copy()
        sp1 = RT1.detector_spectrum(0)
        src1 = RT1.source_spectrum(0)
    # the job: every rank traces its shard of the same bundle, images / messages / extents are reduced
    RT = scenes.SCENES[name](ot)
    RT.trace_rays(p0, s0, pol0, w0, wl, hurb_z=hz, N_list=g["N_list"], sharded=True)
    b, e = dist.shard_range(N)
    assert RT.rays.N == e - b and RT.rays.ray_begin == b and RT.rays.N_global == N
    im = RT.detector_image()
    cnt, dat = im.counts, im.data
    assert np.array_equal(RT._msgs, msgs1), (name, RT._msgs, msgs1)
    assert np.array_equal(np.array(im.extent), ext1), (name, im.extent, ext1)
    assert np.array_equal(cnt, cnt1), (name, int(np.abs(cnt.astype(np.int64) - cnt1).sum()))
    scale = np.abs(dat1).max(axis=(0, 1))
    assert np.all(np.abs(dat - dat1) <= 1e-12*np.abs(dat1) + 1e-15*scale), name       # summation order only
    # the shard's rows of the ray storage are the rows of the single-GPU storage
    assert np.array_equal(RT.rays.p_list, RT1.rays.p_list[b:e]) and np.array_equal(RT.rays.w_list, RT1.rays.w_list[b:e])
    # spectra: counts, range and bin sums reduced over the ranks
    sp = RT.detector_spectrum(0)
    assert np.array_equal(sp._wls, sp1._wls) and np.allclose(sp._vals, sp1._vals, rtol=1e-6, atol=1e-12*sp1._vals.max())
    src = RT.source_spectrum(0)
    assert np.array_equal(src._wls, src1._wls) and np.allclose(src._vals, src1._vals, rtol=1e-6, atol=1e-12*src1._vals.max())
    report[name] = dict(N=int(N), hits=int(cnt.sum()), power=float(im.power()), msgs=[int(v) for v in RT._msgs.sum(axis=1)])

# focus search restricted to one source: with two equal sources on two ranks one rank holds no ray of the source
g = fixture("spherical_aberration")
p0, s0, pol0, w0, wl, hz = gu.bundle(g)
with dist.local_mode():
    RT1 = scenes.spherical_aberration(ot)
    RT1.trace_rays(p0, s0, pol0, w0, wl, N_list=g["N_list"])
    ref = [RT1.focus_search("RMS Spot Size", 20.0, source_index=k)[0].x for k in (0, 1)]
RT = scenes.spherical_aberration(ot)
RT.trace_rays(p0, s0, pol0, w0, wl, N_list=g["N_list"], sharded=True)
for k in (0, 1):
    res, info = RT.focus_search("RMS Spot Size", 20.0, source_index=k)
    assert abs(res.x - ref[k]) <= 1e-9*abs(ref[k]), (k, res.x, ref[k])
res, info = RT.focus_search("Irradiance Variance", 20.0, source_index=1)
report["focus"] = dict(rms=[float(v) for v in ref], irr_var=float(res.x))

# a status bit raised on one rank only reaches every rank (no hang): index below one on the shard of rank 0
RT = ot.Raytracer(outline=[-5, 5, -5, 5, -5, 10])
RT.add(ot.RaySource(ot.CircularSurface(r=1), pos=[0, 0, -1], spectrum=ot.LightSpectrum("Rectangle", wl0=400, wl1=700)))
wls = np.array([380.0, 540.0, 560.0, 780.0])
RT.add(ot.Lens(ot.CircularSurface(r=3), ot.CircularSurface(r=3), pos=[0, 0, 0], d=0.5,
               n=ot.RefractionIndex("Data", wls=wls, vals=np.array([1.5, 1.5, 1.5, 1.5]))))
Nn = 2000
pp = np.zeros((Nn, 3)); pp[:, 2] = -1.0
ss = np.zeros((Nn, 3)); ss[:, 2] = 1.0
pl = np.zeros((Nn, 3), dtype=np.float32); pl[:, 0] = 1
wln = np.full(Nn, 550.0, dtype=np.float32)
wln[:10] = 379.0          # outside the table: np.interp(..., left=0) -> n = 0 < 1, only in rank 0's shard
raised = False
try:
    RT.trace_rays(pp, ss, pl, np.full(Nn, 1/Nn, dtype=np.float32), wln, sharded=True)
except RuntimeError:
    raised = True
flags = torch.tensor([int(raised)], device="cuda")
td.all_reduce(flags)
assert int(flags.item()) == world, "every rank must raise"
report["status_or"] = True

td.barrier()
if rank == 0:
    out = pathlib.Path(os.environ.get("MGPU_REPORT", "/tmp/mgpu_report.json"))
    out.write_text(json.dumps(dict(world=world, **report)))
    print("mgpu worker ok", json.dumps(report))
td.destroy_process_group()
