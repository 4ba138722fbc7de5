"""Thin Python front of the CUDA engine: torch provides device memory, streams and (elsewhere)
torch.distributed; every computation is a call through the C ABI of include/otb.h (ctypes, _cabi.py).

There is no CPU fallback here: without a CUDA device `ensure_init()` raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import check

_initialised = None
_inited_libs = set()


def init_lib(l):
    """otb_init once per loaded engine variant and device (cudaGetDeviceProperties is not free)"""
    key = (id(l), _torch().cuda.current_device())
    if key not in _inited_libs:
        check(l.otb_init(key[1]), l)
        _inited_libs.add(key)
    return l


def _torch():
    import torch
    return torch


def ensure_init():
    """Loads libotb.so and binds it to the current CUDA device (torch.cuda.current_device())."""
    global _initialised
    torch = _torch()
    if not torch.cuda.is_available():
        raise _cabi.EngineError("No CUDA device available: the optrace_b200 engine runs on B200 GPUs only "
                                "(there is no CPU fallback).")
    dev = torch.cuda.current_device()
    if _initialised != dev:
        l = _cabi.lib()
        check(l.otb_init(dev), l)
        _initialised = dev
    return _cabi.lib()


def stream_ptr():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


_side = {}


def side_stream():
    """per-device side stream: image all-reduces and device->host downloads run here, behind an event of the
    compute stream, so that the next trace overlaps them"""
    torch = _torch()
    d = torch.cuda.current_device()
    if d not in _side:
        _side[d] = torch.cuda.Stream(device=d)
    return _side[d]


_gen = {}


def gen_stream():
    """per-device stream of the ray generator: the bundle of the next trace does not depend on the detector kernels
    of the previous one that may still be queued on the compute stream, so it is generated beside them"""
    torch = _torch()
    d = torch.cuda.current_device()
    if d not in _gen:
        _gen[d] = torch.cuda.Stream(device=d)
    return _gen[d]


_pinned_pool = {}


def pinned_take(shape, dtype):
    """pinned host buffer from a small free list (page-locking 143 MB costs tens of ms)"""
    torch = _torch()
    key = (tuple(shape), dtype)
    free = _pinned_pool.setdefault(key, [])
    return free.pop() if free else torch.empty(shape, dtype=dtype, pin_memory=True)


def pinned_give(buf):
    free = _pinned_pool.setdefault((tuple(buf.shape), buf.dtype), [])
    if len(free) < 3:
        free.append(buf)


def device():
    torch = _torch()
    return torch.device("cuda", torch.cuda.current_device())


def dptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def to_dev(a: np.ndarray, dtype=None):
    """host array -> flat device tensor in Fortran (plane) order"""
    torch = _torch()
    a = np.asarray(a, dtype=dtype)
    flat = np.ascontiguousarray(a.reshape(-1, order="F"))
    return torch.from_numpy(flat).to(device(), non_blocking=False)


# ---------------------------------------------------------------------------------------------------
# scene handle
# ---------------------------------------------------------------------------------------------------
class SceneHandle:
    """Device-resident copy of a FlatScene (otb_scene_create / otb_scene_destroy)."""

    def __init__(self, flat, specialised="cached"):
        """specialised: True = build (or load) the scene-specialised engine variant (specialise.py, ~40 s of nvcc
        on a cache miss); "cached" = use it only if it is already in the in-tree cache; False = generic kernels."""
        self.specialised = False
        self.lib = None
        if specialised:
            from . import specialise, userfunc
            uh = userfunc.generate_header(flat.user_funcs) if flat.user_funcs else None
            path = specialise.cached_library(flat, uh) if specialised == "cached" else \
                specialise.build_specialised_library(flat, uh)
            if path is not None:
                ensure_init()
                self.lib = init_lib(_cabi.lib(path))
                self.specialised = True
        if self.lib is None:
            self.lib = ensure_init() if not flat.user_funcs else _user_lib(flat)
        self.flat = flat
        desc = flat.to_ctypes()
        h = C.c_void_p()
        check(self.lib.otb_scene_create(C.byref(desc), C.byref(h)), self.lib)
        self.handle = h
        self.nt = flat.nt

    def reupload(self):
        """send the (unchanged) flattened scene host -> device again: otb_scene_update, asynchronous, no
        allocation (a cudaFree here would synchronise the device and serialise the overlapped image download)"""
        self._desc = self.flat.to_ctypes()          # kept alive until the asynchronous copy has been consumed
        check(self.lib.otb_scene_update(self.handle, C.byref(self._desc), stream_ptr()), self.lib)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.otb_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _user_lib(flat):
    """engine build specialised with the scene's user callables (userfunc.py)"""
    from . import userfunc
    ensure_init()
    path = userfunc.build_specialised_library(flat.user_funcs)
    return init_lib(_cabi.lib(path))


# ---------------------------------------------------------------------------------------------------
# trace
# ---------------------------------------------------------------------------------------------------
class DeviceRays:
    """initial bundle on the device (SoA planes)"""

    def __init__(self, N, p0, s0, pol0, w0, wl, hurb_z=None, seed=0, ray_offset=0):
        self.N, self.p0, self.s0, self.pol0, self.w0, self.wl = N, p0, s0, pol0, w0, wl
        self.hurb_z, self.seed, self.ray_offset = hurb_z, seed, ray_offset

    @staticmethod
    def from_host(p, s, pol, w, wl, hurb_z=None, seed=0):
        N = p.shape[0]
        return DeviceRays(N, to_dev(p, np.float64), to_dev(s, np.float64),
                          None if pol is None else to_dev(pol, np.float32),
                          to_dev(w, np.float32), to_dev(wl, np.float32),
                          None if hurb_z is None else _torch().from_numpy(
                              np.ascontiguousarray(hurb_z, dtype=np.float64).ravel()).to(device()), seed)

    @staticmethod
    def generated(N, sources, n_sources, aux_d, seed, ray_offset):
        """bundle that exists only as a description: the trace kernels draw the rays themselves (OtbRays.gen_h)"""
        r = DeviceRays(N, None, None, None, None, None, None, seed, ray_offset)
        g = _cabi.OtbGenerator()
        g.sources_h = C.cast(sources, C.POINTER(_cabi.OtbSource))
        g.n_sources = n_sources
        g.gen_aux_d = aux_d.data_ptr()
        r._gen = (g, sources, aux_d)         # keeps the ctypes records and the table buffer alive
        return r

    def c_struct(self):
        r = _cabi.OtbRays()
        r.N = self.N
        r.p0_d, r.s0_d, r.pol0_d = dptr(self.p0), dptr(self.s0), dptr(self.pol0)
        r.w0_d, r.wl_d, r.hurb_z_d = dptr(self.w0), dptr(self.wl), dptr(self.hurb_z)
        r.seed, r.ray_offset = self.seed, self.ray_offset
        gen = getattr(self, "_gen", None)
        r.gen_h = C.cast(C.pointer(gen[0]), C.c_void_p) if gen is not None else None
        return r


class DeviceStore:
    """per-surface ray storage on the device, byte layout of RayStorage (ray_storage.py:80-90)"""

    def __init__(self, N: int, nt: int, no_pol: bool):
        torch = _torch()
        d = device()
        self.N, self.nt, self.no_pol = N, nt, no_pol
        self.p = torch.empty(N*nt*3, dtype=torch.float64, device=d)
        self.s = torch.empty(N*3, dtype=torch.float64, device=d)
        self.w = torch.empty(N*nt, dtype=torch.float32, device=d)
        self.n = torch.empty(N*nt, dtype=torch.float64, device=d)
        self.wl = torch.empty(N, dtype=torch.float32, device=d)
        self.pol = None if no_pol else torch.empty(N*nt*3, dtype=torch.float32, device=d)
        self.status = None      # status tensor of the trace that filled the store

    def c_struct(self):
        s = _cabi.OtbRayStore()
        s.N, s.nt = self.N, self.nt
        s.p_d, s.s_d, s.pol_d = dptr(self.p), dptr(self.s), dptr(self.pol)
        s.w_d, s.n_d, s.wl_d = dptr(self.w), dptr(self.n), dptr(self.wl)
        s.trace_status_d = dptr(self.status)
        return s

    @staticmethod
    def nbytes(N: int, nt: int, no_pol: bool) -> int:
        """RayStorage.storage_size (ray_storage.py:92-104)"""
        fpol = 4*N*nt*3 if not no_pol else 8
        return N*nt*3*8 + N*3*8 + fpol + N*nt*4 + N*nt*8 + N*4


def trace_store(scene: SceneHandle, rays: DeviceRays, store: DeviceStore | None = None, msgs=None, sync=True,
                events=None):
    """otb_trace_store: returns (store, msgs int64 tensor (5, nt) on device, status tensor).
    `events` = (start, stop) torch.cuda.Event pair recorded tightly around the kernel launch."""
    torch = _torch()
    if store is None:
        store = DeviceStore(rays.N, scene.nt, scene.flat.no_pol)
    if msgs is None:
        msgs = torch.zeros(_cabi.NMSG*scene.nt, dtype=torch.int64, device=device())
    status = torch.zeros(1, dtype=torch.int32, device=device())
    r, s = rays.c_struct(), store.c_struct()
    if events is not None:
        events[0].record()
    rc = scene.lib.otb_trace_store(scene.handle, C.byref(r), C.byref(s), dptr(msgs), dptr(status), stream_ptr())
    if events is not None:
        events[1].record()
    check(rc, scene.lib)
    store.status = status        # the detector search reads OTB_STATUS_Z_DECREASE from it on the device
    if sync:
        raise_status(int(status.item()))
    return store, msgs.view(_cabi.NMSG, scene.nt), status


def trace_render(scene: SceneHandle, rays: DeviceRays, det_recs: list, extents=None, grids=None, imgs=None,
                 cnts=None, msgs=None, status=None):
    """otb_trace_render (fused trace + detector + binning, no per-surface storage).

    Bin mode: `extents` (list of [x0,x1,y0,y1]), `grids` (list of (Nx, Ny)), `imgs`/`cnts` (device tensors,
    accumulated in place).  Range mode (imgs is None): returns a (n_det, 4) device tensor with the hit ranges.
    At most 8 detectors per launch.  `status`: device int32[1] the caller checks later (raise_status) — bin mode then
    returns without synchronising the host."""
    torch = _torch()
    deferred = status is not None
    n = len(det_recs)
    dets = (_cabi.OtbDetector*n)()
    from .scene import fill_detector
    for k, rec in enumerate(det_recs):
        fill_detector(dets[k], rec)
    if msgs is None:
        msgs = torch.zeros(_cabi.NMSG*scene.nt, dtype=torch.int64, device=device())
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=device())
    r = rays.c_struct()
    if imgs is None:
        rng = torch.tensor([np.inf, -np.inf, np.inf, -np.inf]*n, dtype=torch.float64, device=device())
        check(scene.lib.otb_trace_render(scene.handle, C.byref(r), n, dets, None, None, None, None, None, dptr(rng),
                                         dptr(msgs), dptr(status), stream_ptr()), scene.lib)
        if not deferred:
            raise_status(int(status.item()))
        return rng.view(n, 4)
    ext = (C.c_double*(4*n))(*[float(v) for e in extents for v in e])
    nx = (C.c_int32*n)(*[int(g[0]) for g in grids])
    ny = (C.c_int32*n)(*[int(g[1]) for g in grids])
    ip = (C.c_void_p*n)(*[t.data_ptr() for t in imgs])
    cp = (C.c_void_p*n)(*[t.data_ptr() for t in cnts]) if cnts is not None else None
    check(scene.lib.otb_trace_render(scene.handle, C.byref(r), n, dets, ext, nx, ny, ip, cp, None,
                                     dptr(msgs), dptr(status), stream_ptr()), scene.lib)
    if not deferred:
        raise_status(int(status.item()))
    return msgs.view(_cabi.NMSG, scene.nt)


def hurb_normals(lib, N: int, seed: int, ray_offset: int, n_slots: int) -> np.ndarray:
    """(n_slots, 2, N) standard normal deviates the trace kernels draw for the HURB apertures of a device-RNG trace
    (otb_hurb_normals): lets the oracle replay exactly that trace"""
    torch = _torch()
    out = torch.empty((max(n_slots, 1), 2, max(N, 1)), dtype=torch.float64, device=device())
    for k in range(n_slots):
        check(lib.otb_hurb_normals(N, seed, ray_offset, k, dptr(out[k, 0]), dptr(out[k, 1]), stream_ptr()), lib)
    return out[:n_slots, :, :N].cpu().numpy()


def raise_status(st: int):
    if st & 8:
        raise RuntimeError("All ray divergences s need to be in positive z-divergence")
    if st & 1:
        raise TimeoutError("Timeout after 200 iterations in hit finding. Try retracing.")
    if st & 2:
        raise RuntimeError("Refraction index below 1 for a traced wavelength.")
    if st & 4:
        raise _cabi.EngineError("unsupported feature reached on the device")


# ---------------------------------------------------------------------------------------------------
# detector
# ---------------------------------------------------------------------------------------------------
_det_meta_template = {}


def _det_struct(rec: dict) -> _cabi.OtbDetector:
    from .scene import fill_detector
    d = _cabi.OtbDetector()
    fill_detector(d, rec)
    return d


def detector_hits(lib, store: DeviceStore, det_rec: dict, ray_begin: int = 0, ray_end: int | None = None):
    """otb_detector_hits: returns (hx, hy, hw device tensors over the ray range, range tensor[4], ill count)"""
    torch = _torch()
    ray_end = store.N if ray_end is None else ray_end
    n = ray_end - ray_begin
    d = device()
    hx = torch.empty(max(n, 1), dtype=torch.float64, device=d)
    hy = torch.empty(max(n, 1), dtype=torch.float64, device=d)
    hw = torch.empty(max(n, 1), dtype=torch.float32, device=d)
    # one 48-byte scratch record [range(4) f64 | ill i64 | status i32] so that the caller reads everything back
    # with a single device -> host copy; initialised from a cached device template (no host -> device copy)
    key = d.index
    if key not in _det_meta_template:
        t = torch.zeros(6, dtype=torch.float64)
        t[:4] = torch.tensor([np.inf, -np.inf, np.inf, -np.inf], dtype=torch.float64)
        _det_meta_template[key] = t.to(d)
    meta = _det_meta_template[key].clone()
    rng, ill, status = meta[:4], meta[4:5].view(torch.int64), meta[5:6].view(torch.int32)[:1]
    s = store.c_struct()
    det = _det_struct(det_rec)
    check(lib.otb_detector_hits(C.byref(s), ray_begin, ray_end, C.byref(det), dptr(hx), dptr(hy), dptr(hw),
                                dptr(rng), dptr(ill), dptr(status), stream_ptr()), lib)
    return hx[:n], hy[:n], hw[:n], rng, ill, status, meta


def read_det_meta(meta):
    """(range ndarray[4], ill count, status) from the scratch records of detector_hits of ALL ranks: one all-gather
    of the 48-byte record (none on a single GPU), one synchronising copy, reduced on the host — hit range MIN / MAX
    (auto extent, raytracer.py:1042-1046), ill-conditioned count SUM, status bits OR"""
    from . import dist
    rows = dist.gather_rows(meta)                   # (world, 6) float64
    r = np.array([rows[:, 0].min(), rows[:, 1].max(), rows[:, 2].min(), rows[:, 3].max()])
    ill = int(rows[:, 4].copy().view(np.int64).sum())
    st = 0
    for v in rows[:, 5].copy().view(np.int32).reshape(-1, 2)[:, 0]:
        st |= int(v)
    return r, ill, st


def render_xyzw(lib, x, y, w, wl, extent, Nx: int, Ny: int, img=None, cnt=None):
    """otb_render_xyzw on device tensors; returns (img (Ny,Nx,4) f64, cnt (Ny,Nx) i32)"""
    torch = _torch()
    d = device()
    if img is None:
        img = torch.zeros((Ny, Nx, 4), dtype=torch.float64, device=d)
    if cnt is None:
        cnt = torch.zeros((Ny, Nx), dtype=torch.int32, device=d)
    e = (C.c_double*4)(*[float(v) for v in extent])
    M = int(x.shape[0])
    check(lib.otb_render_xyzw(dptr(x), dptr(y), dptr(w), dptr(wl), M, e, Nx, Ny, dptr(img), dptr(cnt), stream_ptr()),
          lib)
    return img, cnt


def render_xyzw_host(p, w, wl, extent, Nx: int, Ny: int):
    """RenderImage.render on host arrays: copies to the device, bins there"""
    lib = ensure_init()
    torch = _torch()
    if p is None or not p.shape[0]:
        d = device()
        return (torch.zeros((Ny, Nx, 4), dtype=torch.float64, device=d),
                torch.zeros((Ny, Nx), dtype=torch.int32, device=d))
    x = to_dev(np.ascontiguousarray(p[:, 0]), np.float64)
    y = to_dev(np.ascontiguousarray(p[:, 1]), np.float64)
    return render_xyzw(lib, x, y, to_dev(w, np.float32), to_dev(wl, np.float32), extent, Nx, Ny)


# ---------------------------------------------------------------------------------------------------
# stand-alone array evaluation (Surface.find_hit / normals / values, RefractionIndex.__call__)
# ---------------------------------------------------------------------------------------------------
def _surface_args(surf):
    from .scene import standalone_surface, fill_surface
    rec, aux, funcs = standalone_surface(surf)
    lib = ensure_init()
    if funcs:
        from . import userfunc
        lib = init_lib(_cabi.lib(userfunc.build_specialised_library(funcs, api_only=True)))
    S = _cabi.OtbSurface()
    fill_surface(S, rec)
    aux = np.ascontiguousarray(aux if aux.shape[0] else np.zeros(1), dtype=np.float64)
    return lib, S, aux


def surface_find_hit(surf, p: np.ndarray, s: np.ndarray):
    torch = _torch()
    lib, S, aux = _surface_args(surf)
    N = p.shape[0]
    pd, sd = to_dev(p, np.float64), to_dev(s, np.float64)
    ph = torch.empty(3*max(N, 1), dtype=torch.float64, device=device())
    hit = torch.empty(max(N, 1), dtype=torch.uint8, device=device())
    ill = torch.empty(max(N, 1), dtype=torch.uint8, device=device())
    check(lib.otb_surface_find_hit(C.byref(S), aux.ctypes.data_as(C.POINTER(C.c_double)), aux.shape[0], N,
                                   dptr(pd), dptr(sd), dptr(ph), dptr(hit), dptr(ill), stream_ptr()), lib)
    ph = ph[:3*N].cpu().numpy().reshape((N, 3), order="F")
    return ph, hit[:N].cpu().numpy().astype(bool), ill[:N].cpu().numpy().astype(bool)


def surface_normals(surf, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    torch = _torch()
    lib, S, aux = _surface_args(surf)
    N = x.shape[0]
    xd, yd = to_dev(x, np.float64), to_dev(y, np.float64)
    n = torch.empty(3*max(N, 1), dtype=torch.float64, device=device())
    check(lib.otb_surface_normals(C.byref(S), aux.ctypes.data_as(C.POINTER(C.c_double)), aux.shape[0], N,
                                  dptr(xd), dptr(yd), dptr(n), stream_ptr()), lib)
    return n[:3*N].cpu().numpy().reshape((N, 3), order="F")


def surface_values(surf, x: np.ndarray, y: np.ndarray):
    """(values, mask) evaluated on the device"""
    torch = _torch()
    lib, S, aux = _surface_args(surf)
    N = x.shape[0]
    xd, yd = to_dev(x, np.float64), to_dev(y, np.float64)
    z = torch.empty(max(N, 1), dtype=torch.float64, device=device())
    m = torch.empty(max(N, 1), dtype=torch.uint8, device=device())
    check(lib.otb_surface_values(C.byref(S), aux.ctypes.data_as(C.POINTER(C.c_double)), aux.shape[0], N,
                                 dptr(xd), dptr(yd), dptr(z), dptr(m), stream_ptr()), lib)
    return z[:N].cpu().numpy(), m[:N].cpu().numpy().astype(bool)


def medium_eval(ri, wl: np.ndarray) -> np.ndarray:
    from .scene import FlatScene, fill_medium
    torch = _torch()
    fs = FlatScene()
    fs._media_objs = []
    fs.add_medium(ri)
    lib = ensure_init() if not fs.user_funcs else _user_lib(fs)
    M = _cabi.OtbMedium()
    fill_medium(M, fs.media[0])
    aux = np.ascontiguousarray(fs.aux if fs.aux.shape[0] else np.zeros(1), dtype=np.float64)
    N = wl.shape[0]
    wd = to_dev(wl, np.float64)
    n = torch.empty(max(N, 1), dtype=torch.float64, device=device())
    check(lib.otb_medium_eval(C.byref(M), aux.ctypes.data_as(C.POINTER(C.c_double)), aux.shape[0], N,
                              dptr(wd), dptr(n), stream_ptr()), lib)
    return n[:N].cpu().numpy()


def sphere_projection(surf, p: np.ndarray, method: str) -> np.ndarray:
    from .scene import detector_record, fill_detector
    torch = _torch()
    lib = ensure_init()
    rec = detector_record(surf, method, None)
    if method == "Orthographic":
        return p.copy()
    det = _cabi.OtbDetector()
    fill_detector(det, rec)
    N = p.shape[0]
    pd = to_dev(p, np.float64)
    out = torch.empty(3*max(N, 1), dtype=torch.float64, device=device())
    check(lib.otb_sphere_projection(C.byref(det.surface), rec["projection"], N, dptr(pd), dptr(out), stream_ptr()), lib)
    return out[:3*N].cpu().numpy().reshape((N, 3), order="F")


# ---------------------------------------------------------------------------------------------------
# image post-processing (RenderImage.get, render_image.py:131-222)
# ---------------------------------------------------------------------------------------------------
def image_get(data_dev, fact: int, mode: int, scale: float, L_th: float, chroma_scale):
    """join-bins rescale + conversion of a device (Ny, Nx, 4) histogram; returns the host array of the result"""
    torch = _torch()
    lib = ensure_init()
    Ny, Nx, _ = data_dev.shape
    st = stream_ptr()
    if fact != 1:
        img = torch.empty((Ny//fact, Nx//fact, 4), dtype=torch.float64, device=data_dev.device)
        check(lib.otb_image_rescale(dptr(data_dev), Ny, Nx, fact, dptr(img), st), lib)
    else:
        img = data_dev
    H, W = int(img.shape[0]), int(img.shape[1])
    npx = H*W
    stats = torch.empty(8, dtype=torch.float64, device=data_dev.device)
    cs = -1.0
    if mode >= 2:           # every mode but irradiance / illuminance normalises by image-wide extrema
        check(lib.otb_image_stats(dptr(img), npx, 1, 0.0, dptr(stats), st), lib)
    if mode == 3:           # perceptual intent: srgb.py:303-352
        s1 = stats.cpu().numpy()
        if s1[1] != 0.0 or chroma_scale is not None:
            check(lib.otb_image_stats(dptr(img), npx, 2, L_th, dptr(stats), st), lib)
            cmin = float(stats[6].item())
            cr = np.sqrt(cmin) if np.isfinite(cmin) else 1.0
            fact_c = float(np.clip(cr, 0.32, 1.0))
            cs = float(chroma_scale) if chroma_scale is not None else fact_c
            check(lib.otb_image_stats(dptr(img), npx, 3, cs, dptr(stats), st), lib)
    out = torch.empty((H, W, 3) if mode in (2, 3) else (H, W), dtype=torch.float64, device=data_dev.device)
    check(lib.otb_image_convert(dptr(img), npx, mode, float(scale), cs, dptr(stats), dptr(out), st), lib)
    return out.cpu().numpy()


def image_convolve(data_dev, psf: np.ndarray):
    """otb_image_convolve: zero-padded "same" convolution of every channel of a device (Ny, Nx, 4) image with a
    host (K, K) kernel, negatives removed; returns a new device tensor"""
    torch = _torch()
    lib = ensure_init()
    Ny, Nx, _ = data_dev.shape
    psf_d = torch.from_numpy(np.ascontiguousarray(psf, dtype=np.float64)).to(data_dev.device)
    out = torch.empty_like(data_dev)
    check(lib.otb_image_convolve(dptr(data_dev), Ny, Nx, dptr(psf_d), int(psf.shape[0]), dptr(out), stream_ptr()), lib)
    return out


# ---------------------------------------------------------------------------------------------------
# sparse transport of detector images (otb_tiles.cu)
# ---------------------------------------------------------------------------------------------------
TILE = 32
_tile_cap = {}        # (Ny, Nx) -> tile capacity learnt from earlier images of that shape
_tile_last = {}       # (Ny, Nx) -> (event, header tensor) of the latest pack: read without waiting once it is done


def tile_capacity(shape) -> int:
    """capacity (tiles) for the next packed image of this shape: twice what the previous one needed (power of two,
    so that the pinned staging buffers are reused), an eighth of the image to start with; 0 = dense transport"""
    Ny, Nx = int(shape[0]), int(shape[1])
    total = -(-Ny//TILE)*(-(-Nx//TILE))
    key = (Ny, Nx)
    last = _tile_last.get(key)
    if last is not None and last[0].query():          # the pinned copy of the header has landed: no synchronisation
        _tile_cap[key] = int(last[1][0])
        _tile_last.pop(key, None)
    need = _tile_cap.get(key)
    cap = max(64, total//8) if need is None else max(64, 2*need)
    cap = 1 << (cap - 1).bit_length()
    return cap if cap*2 <= total else 0


class TilePack:
    """occupied T x T tiles of a device image (Ny, Nx, 4): mask -> [union over ranks] -> pack -> [sum over ranks] ->
    unpack.  All calls enqueue on the current stream."""

    def __init__(self, lib, img, cap: int):
        torch = _torch()
        self.lib, self.img, self.cap = lib, img, cap
        self.Ny, self.Nx = int(img.shape[0]), int(img.shape[1])
        self.ntiles = -(-self.Ny//TILE)*(-(-self.Nx//TILE))
        d = img.device
        self.mask = torch.zeros(self.ntiles, dtype=torch.int32, device=d)
        self.header = torch.zeros(2 + cap, dtype=torch.int32, device=d)
        self.packed = torch.empty(cap*TILE*TILE*4, dtype=torch.float64, device=d)
        self.reduced = False      # the packed tiles hold the SUM over all ranks (dist.allreduce_image_async)

    def make_mask(self):
        check(self.lib.otb_image_tiles_mask(dptr(self.img), self.Ny, self.Nx, TILE, dptr(self.mask), stream_ptr()), self.lib)

    def pack(self):
        check(self.lib.otb_image_tiles_pack(dptr(self.img), self.Ny, self.Nx, TILE, dptr(self.mask), self.cap,
                                            dptr(self.header), dptr(self.packed), stream_ptr()), self.lib)

    def unpack(self):
        check(self.lib.otb_image_tiles_unpack(dptr(self.img), self.Ny, self.Nx, TILE, dptr(self.header), self.cap,
                                              dptr(self.packed), stream_ptr()), self.lib)

    def remember(self):
        """asynchronous pinned copy of [count, overflow] on the current stream: the next image of this shape sizes its
        capacity from it (tile_capacity) without any host synchronisation"""
        torch = _torch()
        h = pinned_take((2,), torch.int32)
        h.copy_(self.header[:2], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _tile_last[(self.Ny, self.Nx)] = (ev, h)


_dense_pool = {}      # (Ny, Nx) -> list of (padded zero array with the tiles of its last use cleared again)


def assemble_tiles(shape, header: np.ndarray, packed: np.ndarray):
    """dense host image from the packed tiles (header = [count, overflow, ids...]): one scatter of all tiles into a
    zero image padded to whole tiles.  The padded arrays are pooled: a fresh 143 MB np.zeros costs a page fault per
    4 KB page the tiles touch (~1 ms), a recycled one (release_dense clears exactly the tiles that were written) none.
    Returns (the (Ny, Nx, 4) view, token for release_dense)."""
    Ny, Nx = int(shape[0]), int(shape[1])
    nty, ntx = -(-Ny//TILE), -(-Nx//TILE)
    n = int(header[0])
    free = _dense_pool.setdefault(("padded", nty*TILE, ntx*TILE), [])
    out = free.pop() if free else np.zeros((nty*TILE, ntx*TILE, 4), dtype=np.float64)
    ids = header[2:2 + n].astype(np.int64)
    if n:
        tiles = packed.reshape(-1, TILE, TILE, 4)[:n]
        out.reshape(nty, TILE, ntx, TILE, 4)[ids//ntx, :, ids % ntx] = tiles
    return out[:Ny, :Nx], (out, ids)


_assembler = None


def assemble_submit(shape, ready_event, hdr_t, buf_t):
    """assemble_tiles on the worker thread once `ready_event` (the pinned copy) has completed.  Returns a future of
    (dense view, pool token), or of None when the tile capacity overflowed (the caller then takes the dense path)."""
    global _assembler
    if _assembler is None:
        import concurrent.futures
        _assembler = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="otb-assemble")

    def task():
        ready_event.synchronize()           # releases the GIL while waiting
        h = hdr_t.numpy()
        if int(h[1]):
            return None
        return assemble_tiles(shape, h, buf_t.numpy())

    return _assembler.submit(task)


def release_dense_async(token) -> None:
    """release_dense on the assembler thread (a RenderImage usually dies on the caller's critical path, right before
    the next trace is queued); falls back to the caller's thread when the worker does not exist (yet)"""
    if _assembler is None:
        release_dense(token)
        return
    try:
        _assembler.submit(release_dense, token)
    except RuntimeError:          # interpreter shutting down
        release_dense(token)


def release_dense(token) -> None:
    """give a padded host image back to the pool (called when its RenderImage dies): the written tiles are zeroed"""
    out, ids = token
    nty, ntx = out.shape[0]//TILE, out.shape[1]//TILE
    if ids.shape[0]:
        out.reshape(nty, TILE, ntx, TILE, 4)[ids//ntx, :, ids % ntx] = 0.0
    free = _dense_pool.setdefault(("padded", out.shape[0], out.shape[1]), [])
    if len(free) < 3:
        free.append(out)


# ---------------------------------------------------------------------------------------------------
# spectrum histograms (LightSpectrum.render, light_spectrum.py:40-79)
# ---------------------------------------------------------------------------------------------------
def spectrum_histogram(lib, wl, w, positive_only: bool, wavelength_range):
    """(vals float64[N], edges float32[N + 1]) of the weighted wavelength histogram of device tensors wl, w
    (float32).  Bin count, range and float32 edges follow LightSpectrum.render / np.histogram; the counts and the
    range of all ranks are combined with one all-gather, the bin sums with one all-reduce."""
    torch = _torch()
    from . import dist
    d = device()
    M = int(wl.shape[0])
    rec = torch.zeros(4, dtype=torch.int64, device=d)         # [used, non-zero, (min, max) as two float32]
    rng = rec[2:3].view(torch.float32)
    rng.copy_(torch.tensor([np.inf, -np.inf], dtype=torch.float32))
    check(lib.otb_spectrum_stats(dptr(wl), dptr(w), M, int(positive_only), dptr(rec), dptr(rng), stream_ptr()), lib)
    rows = dist.gather_rows(rec)                               # one collective, one host synchronisation
    used, nz = int(rows[:, 0].sum()), int(rows[:, 1].sum())
    fr = rows[:, 2].copy().view(np.float32).reshape(-1, 2)
    N = max(51, np.sqrt(nz)/2)
    N = 1 + 2*(int(N)//2)
    if not used:
        return None, N
    wl0, wl1 = np.float32(fr[:, 0].min()), np.float32(fr[:, 1].max())
    if np.abs(wl0 - wl1) < 1:
        wl0, wl1 = max(wl0 - 1, wavelength_range[0]), min(wl0 + 1, wavelength_range[1])
    # the edges numpy itself would use (np.histogram -> _get_bin_edges -> np.linspace in the promoted dtype)
    edges = np.histogram_bin_edges(np.array([wl0, wl1], dtype=np.float32), bins=N, range=[wl0, wl1])
    edges32 = np.ascontiguousarray(edges, dtype=np.float32)
    e_d = torch.from_numpy(edges32).to(d)
    hist = torch.zeros(N, dtype=torch.float64, device=d)
    check(lib.otb_spectrum_hist(dptr(wl), dptr(w), M, int(positive_only), dptr(e_d), N, dptr(hist), stream_ptr()), lib)
    dist.allreduce_sum_(hist)
    return hist.cpu().numpy(), edges


# ---------------------------------------------------------------------------------------------------
# focus search (Raytracer.focus_search, raytracer.py:1354-1640)
# ---------------------------------------------------------------------------------------------------
class FocusLines:
    """device arrays of the auxiliary lines hit(z) = pa + sb z of the selected rays (otb_focus_prepare)"""

    def __init__(self, lib, store: DeviceStore, begin: int, end: int, z: float):
        torch = _torch()
        d = device()
        n = end - begin
        self.lib, self.n = lib, n
        self.pax, self.pay, self.sbx, self.sby = (torch.empty(max(n, 1), dtype=torch.float64, device=d) for _ in range(4))
        self.w = torch.empty(max(n, 1), dtype=torch.float32, device=d)
        self.use = torch.empty(max(n, 1), dtype=torch.uint8, device=d)
        cnt = torch.zeros(1, dtype=torch.int64, device=d)
        s = store.c_struct()
        check(lib.otb_focus_prepare(C.byref(s), begin, end, float(z), dptr(self.pax), dptr(self.pay), dptr(self.sbx),
                                    dptr(self.sby), dptr(self.w), dptr(self.use), dptr(cnt), stream_ptr()), lib)
        from . import dist
        self._dist = dist
        dist.allreduce_sum_(cnt)                     # several GPUs: every rank holds the lines of its shard
        self.n_use = int(cnt.item())
        self._par = torch.zeros(5, dtype=torch.float64, device=d)
        self._out = torch.zeros(4, dtype=torch.float64, device=d)
        self._rng = torch.zeros(4, dtype=torch.float64, device=d)

    def _args(self):
        return (dptr(self.pax), dptr(self.pay), dptr(self.sbx), dptr(self.sby), dptr(self.w), dptr(self.use), self.n)

    def moments(self, mode: int, par) -> np.ndarray:
        torch = _torch()
        p = np.zeros(5)
        p[:len(par)] = par
        self._par.copy_(torch.from_numpy(p))
        check(self.lib.otb_focus_moments(*self._args(), mode, dptr(self._par), dptr(self._out), stream_ptr()), self.lib)
        self._dist.allreduce_sum_(self._out)
        return self._out.cpu().numpy().copy()

    def image(self, z: float, npx: int):
        """(extent [x0, x1, y0, y1], weighted (npx, npx) histogram) of the hit positions at z"""
        torch = _torch()
        img = torch.empty((npx, npx), dtype=torch.float64, device=self.pax.device)
        if self._dist.world() > 1:       # range of all shards first, then every shard bins over the common grid
            check(self.lib.otb_focus_image(*self._args(), float(z), npx, 1, dptr(self._rng), dptr(img), stream_ptr()), self.lib)
            self._dist.allreduce_range_(self._rng)
            check(self.lib.otb_focus_image(*self._args(), float(z), npx, 2, dptr(self._rng), dptr(img), stream_ptr()), self.lib)
            self._dist.allreduce_sum_(img)
        else:
            check(self.lib.otb_focus_image(*self._args(), float(z), npx, 0, dptr(self._rng), dptr(img), stream_ptr()), self.lib)
        return self._rng.cpu().numpy().copy(), img.cpu().numpy()
