"""Multi-GPU plumbing: one process per GPU (torchrun), rays sharded by contiguous global ray-id ranges,
no collective on the propagation path.  Only detector histograms, hit-extent bounds, message counters and
flux statistics are reduced (NCCL all-reduce over NVLink on GPU tensors; gloo on CPU tensors in the tests).

The reference has no distributed backend at all; its analogue is the per-thread ray ranges of
raytracer.py:285-286, 399-405.
"""
from __future__ import annotations

import numpy as np


def _td():
    import torch.distributed as td
    return td


_force_local = False     # tests: run one rank of an initialised job as if it were alone (reference result on one GPU)


class local_mode:
    """context manager: inside, this process behaves like a single-GPU job (no sharding, no collectives)"""

    def __enter__(self):
        global _force_local
        self._old, _force_local = _force_local, True

    def __exit__(self, *a):
        global _force_local
        _force_local = self._old


def is_dist() -> bool:
    if _force_local:
        return False
    td = _td()
    return td.is_available() and td.is_initialized()


def world() -> int:
    return _td().get_world_size() if is_dist() else 1


def rank() -> int:
    return _td().get_rank() if is_dist() else 0


def shard_range(N: int, r: int | None = None, G: int | None = None) -> tuple[int, int]:
    """contiguous global ray-id range [begin, end) of rank r among G ranks; the last rank takes the remainder
    (same rule as RayStorage.thread_rays, ray_storage.py:146-148)"""
    r = rank() if r is None else r
    G = world() if G is None else G
    Np = int(N/G)
    begin = r*Np
    end = begin + Np if r != G - 1 else N
    return begin, end


def source_slices(B_list, begin: int, end: int) -> list[tuple[int, int, int]]:
    """(source index, local start, count) of every source block intersecting [begin, end)
    (the per-thread source walk of ray_storage.py:150-168)"""
    out = []
    for i in range(len(B_list) - 1):
        s, e = max(begin, int(B_list[i])), min(int(B_list[i + 1]), end)
        if e > s:
            out.append((i, s - begin, e - s))
    return out


def shard_sources(N_list, r: int | None = None, G: int | None = None) -> list[tuple[int, int, int]]:
    """Balanced sharding of a generated bundle: rank r takes the r-th slice of EVERY source, as blocks
    (source index, global id of the block's first ray, count) in source order.  Contiguous global ranges
    (shard_range) hand whole sources to single ranks; sources differ in cost — off-axis field points lose most of
    their rays at the stop, on-axis ones none — and the slowest rank sets the pace of every collective (measured on
    the double-Gauss workload: trace kernel 3.92 ms on the rank holding the on-axis source against 3.40 ms average)."""
    r = rank() if r is None else r
    G = world() if G is None else G
    out, B = [], 0
    for i, n in enumerate(int(v) for v in N_list):
        Np = n//G
        b = r*Np
        e = b + Np if r != G - 1 else n
        if e > b:
            out.append((i, B + b, e - b))
        B += n
    return out


def contiguous_blocks(N_list, begin: int, end: int) -> list[tuple[int, int, int]]:
    """the blocks (source index, global first ray id, count) of a contiguous global ray range [begin, end)"""
    B_list = np.concatenate(([0], np.cumsum(np.asarray(N_list, dtype=np.int64))))
    return [(i, begin + start, cnt) for i, start, cnt in source_slices(B_list, begin, end)]


def allreduce_sum_(t):
    """in-place SUM all-reduce (histograms, counters)"""
    if is_dist() and world() > 1:
        td = _td()
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


_image_group = None


def allreduce_sum_async(tensors):
    """SUM all-reduce of GPU tensors on the side stream, ordered after the work queued so far on the current
    stream.  Returns the CUDA event that marks completion (None when there is nothing to reduce): the caller's
    stream goes on with the next trace while NCCL moves the image over NVLink."""
    if not (is_dist() and world() > 1):
        return None
    import torch
    from . import engine
    td = _td()
    global _image_group
    if _image_group is None:
        # own communicator (own NCCL stream): the small synchronous collectives of the next trace (per-source ray
        # counts, messages) must not queue behind a 143 MB image all-reduce.  Created collectively at first use.
        _image_group = td.new_group(backend=td.get_backend())
    side = engine.side_stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for t in tensors:
            td.all_reduce(t, op=td.ReduceOp.SUM, group=_image_group)
            t.record_stream(side)
        ev = torch.cuda.Event()
        ev.record(side)
    return ev


def allreduce_image_async(lib, img):
    """SUM all-reduce of a detector image (Ny, Nx, 4) on the side stream, moving only its occupied tiles
    (engine.TilePack, otb_tiles.cu): MAX all-reduce of the tile mask (the union over the ranks, a few KB), pack by the
    union, SUM all-reduce of the packed tiles, unpack.  Returns (completion event, TilePack or None).  Falls back to
    the dense all-reduce when the image is not sparse; an overflow of the learnt capacity is detected by the consumer
    (RenderImage._wait_device) and repaired with the dense all-reduce."""
    if not (is_dist() and world() > 1):
        return None, None
    import torch
    from . import engine
    td = _td()
    global _image_group
    if _image_group is None:
        _image_group = td.new_group(backend=td.get_backend())
    cap = engine.tile_capacity(img.shape) if img.is_cuda else 0
    if not cap:
        return allreduce_sum_async((img,)), None
    # The packed tiles are a few MB: the collective runs on the COMPUTE stream, right behind the render kernel.  The
    # ranks met in the all-gather of the hit ranges a moment ago, so nobody waits long, and no NCCL blocks sit on
    # the SMs beside the next trace kernel (a one-wave persistent grid: with the all-reduce overlapped on a side
    # stream it ran 3.92 instead of 3.40 ms on 8 GPUs, NCCL's spinning CTAs taking its block slots).
    tp = engine.TilePack(lib, img, cap)
    tp.make_mask()
    td.all_reduce(tp.mask, op=td.ReduceOp.MAX)
    tp.pack()
    td.all_reduce(tp.packed, op=td.ReduceOp.SUM)
    tp.unpack()
    tp.reduced = True
    tp.remember()
    ev = torch.cuda.Event()
    ev.record()
    return ev, tp


def allreduce_range_(rng):
    """rng = [min x, max x, min y, max y] tensor: MIN/MAX all-reduce for the auto extent (raytracer.py:1042-1046)"""
    if is_dist() and world() > 1:
        td = _td()
        sign = rng.new_tensor([-1.0, 1.0, -1.0, 1.0])
        v = rng*sign                    # turn the minima into maxima
        td.all_reduce(v, op=td.ReduceOp.MAX)
        rng.copy_(v*sign)
    return rng


STATUS_BITS = 5      # OTB_STATUS_* bits of include/otb.h


def reduce_msgs_status_begin(msgs, status):
    """Message counters (SUM) and the device status word (bitwise OR) of all ranks with ONE collective: the status
    bits travel bit-expanded behind the counters (NCCL has no bitwise reduction), so every rank raises the same
    exception instead of one rank raising while the others wait in the next collective.  Enqueues the collective
    and an asynchronous copy into pinned host memory on the current stream; reduce_msgs_status_end() waits for it —
    with Raytracer.deferred_status that is after the next synchronisation point, at no extra cost."""
    import torch
    shape = tuple(msgs.shape)
    multi = is_dist() and world() > 1
    if multi:
        td = _td()
        bits = (status.to(torch.int64).reshape(1) >> torch.arange(STATUS_BITS, device=status.device)) & 1
        buf = torch.cat((msgs.reshape(-1).to(torch.int64), bits))
        td.all_reduce(buf, op=td.ReduceOp.SUM)
    else:
        buf = torch.cat((msgs.reshape(-1).to(torch.int64), status.to(torch.int64).reshape(1)))
    if not buf.is_cuda:          # gloo / CPU tensors (tests): nothing to overlap
        return buf, None, shape, multi
    from . import engine
    h = engine.pinned_take(buf.shape, torch.int64)
    h.copy_(buf, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return h, ev, shape, multi


def reduce_msgs_status_end(handle):
    """(msgs ndarray like the input shape, status int) of a reduce_msgs_status_begin handle; one host wait"""
    h, ev, shape, multi = handle
    if ev is not None:
        ev.synchronize()
        a = h.numpy().copy()
        from . import engine
        engine.pinned_give(h)
    else:
        a = h.numpy()
    if multi:
        st = int(sum((1 << i) for i in range(STATUS_BITS) if a[-STATUS_BITS + i] > 0))
        return a[:-STATUS_BITS].reshape(shape).astype(int), st
    return a[:-1].reshape(shape).astype(int), int(a[-1])


def reduce_msgs_status(msgs, status):
    """both halves at once: ONE collective and ONE host synchronisation"""
    return reduce_msgs_status_end(reduce_msgs_status_begin(msgs, status))


def gather_rows(t):
    """small per-rank record (1-D tensor) -> host ndarray (world, len) with one collective and one host
    synchronisation; the caller reduces the rows itself (min / max / sum / or per field)"""
    if is_dist() and world() > 1:
        import torch
        td = _td()
        out = torch.empty((world(), t.numel()), dtype=t.dtype, device=t.device)
        td.all_gather_into_tensor(out, t.reshape(1, -1).contiguous())
        return out.cpu().numpy()
    return t.reshape(1, -1).cpu().numpy()


def allreduce_max_scalar(v: float, device) -> float:
    if is_dist() and world() > 1:
        import torch
        td = _td()
        t = torch.tensor([v], dtype=torch.float64, device=device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())
    return v


def broadcast_ints(a: np.ndarray, device) -> np.ndarray:
    """rank 0's integer array to all ranks (per-source ray counts involve np.random.choice, ray_storage.py:63-68)"""
    if is_dist() and world() > 1:
        import torch
        td = _td()
        t = torch.as_tensor(np.asarray(a, dtype=np.int64), device=device)
        td.broadcast(t, src=0)
        return t.cpu().numpy()
    return np.asarray(a, dtype=np.int64)


_split_cache = None


def shared_split(N: int, powers, device) -> np.ndarray:
    """rays per source, identical on every rank (RayStorage.init, ray_storage.py:56-68).  The split is deterministic
    unless N does not divide by the power ratios: only then the remainder is drawn with np.random.choice and rank
    0's draw is broadcast (a collective plus a host synchronisation that the common case does not pay)."""
    from .ray_storage import split_rays
    global _split_cache
    key = (int(N), tuple(float(p) for p in powers))
    if _split_cache is not None and _split_cache[0] == key:      # deterministic case only (see below)
        return _split_cache[1].copy()
    P = np.asarray(powers, dtype=np.float64)
    if N - int(np.sum((N*P/np.sum(P)).astype(int))) == 0:
        out = split_rays(N, powers)
        _split_cache = (key, out.copy())
        return out
    return broadcast_ints(split_rays(N, powers), device)


def broadcast_floats(a: np.ndarray, device) -> np.ndarray:
    """rank 0's float64 array to all ranks (random sampling positions that every rank has to share)"""
    if is_dist() and world() > 1:
        import torch
        td = _td()
        t = torch.as_tensor(np.asarray(a, dtype=np.float64), device=device)
        td.broadcast(t, src=0)
        return t.cpu().numpy()
    return np.asarray(a, dtype=np.float64)
