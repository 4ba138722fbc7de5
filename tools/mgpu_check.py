"""Multi-GPU consistency check (run under torchrun on N GPUs):
the all-reduced detector image of a sharded trace must bin exactly the rays of all shards, messages must add up,
and every rank must end with the same image.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_check.py"""
import os, sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch, torch.distributed as td
warnings.simplefilter("ignore")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
import optrace_b200 as ot
from optrace_b200 import dist
import scenes
ot.global_options.show_warnings = False
rank, world = dist.rank(), dist.world()
for name in ("double_gauss", "image_render"):
    RT = scenes.SCENES[name](ot)
    N = 1_000_003
    RT.trace(N)
    assert RT.rays.N == sum(c for _, _, c in dist.shard_sources(RT.rays.N_list))
    img = RT.detector_image()
    local_alive = torch.tensor([float(RT.rays.N)], dtype=torch.float64, device="cuda")
    dist.allreduce_sum_(local_alive)
    assert int(local_alive.item()) == N
    cnt = int(img.counts.sum())
    pw = img.power()
    # every rank holds the identical reduced image
    chk = torch.tensor([pw, float(cnt)], dtype=torch.float64, device="cuda")
    mx = chk.clone(); td.all_reduce(mx, op=td.ReduceOp.MAX)
    mn = chk.clone(); td.all_reduce(mn, op=td.ReduceOp.MIN)
    assert torch.equal(mx, mn), (mx, mn)
    # fused path agrees with the store path in total power (same seeds are not shared; statistical agreement)
    RT.ITER_RAYS_STEP = N
    ims = RT.iterative_render(N)
    rel = abs(ims[0].power() - pw)/pw
    if rank == 0:
        print(f"{name}: world={world} N={N} hits={cnt} power={pw:.6f} fused power={ims[0].power():.6f} rel diff {rel:.2e} msgs={RT._msgs.sum(axis=1)}")
    assert rel < 2e-2
# focus search on sharded rays: moments, ranges and cost images are all-reduced, every rank runs the same optimiser
RT = scenes.SCENES["double_gauss"](ot)
RT.trace(1_000_003)
res, info = RT.focus_search("RMS Spot Size", 120.0)
res2, info2 = RT.focus_search("Image Sharpness", 120.0)
chk = torch.tensor([res.x, res.fun, float(info["N"]), res2.x], dtype=torch.float64, device="cuda")
mx = chk.clone(); td.all_reduce(mx, op=td.ReduceOp.MAX)
mn = chk.clone(); td.all_reduce(mn, op=td.ReduceOp.MIN)
assert torch.equal(mx, mn), (mx, mn)
assert 130 < res.x < 150 and info2["bounds"][0] <= res2.x <= info2["bounds"][1], (res.x, res2.x)
if rank == 0:
    print(f"focus_search: world={world} rms focus z={res.x:.6f} (spot {res.fun:.3e} mm, {info['N']} rays), sharpness focus z={res2.x:.4f}")
td.barrier()
if rank == 0:
    print("mgpu_check ok")
td.destroy_process_group()
