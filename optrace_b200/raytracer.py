"""Raytracer of the drop-in API: scene container + trace / detector_image / iterative_render driving the
CUDA engine.  Mirrors the public surface of optrace/tracer/raytracer.py (constructor, class constants, INFOS,
message bookkeeping, error behaviour); the per-ray work of its sub_trace loop, _hit_detector and
RenderImage.render happens in the kernels behind include/otb.h.

Multi-GPU: under torchrun (torch.distributed initialised, one process per GPU) `trace(N)` traces this rank's
contiguous share of the N rays; detector images, hit extents and message counters are all-reduced (dist.py).
"""
from __future__ import annotations

import ctypes as C
from enum import IntEnum

import sys

import numpy as np

from . import _cabi, dist, engine
from .elements import Group, Lens, Filter, Aperture, Detector, RaySource
from .images import RenderImage
from .media import RefractionIndex, LightSpectrum
from .options import global_options, warning
from .ray_storage import RayStorage, split_rays
from .scene import flatten_raytracer, detector_record
from .surfaces import RectangularSurface, RingSurface, SlitSurface, SphericalSurface
from ._state import state_of


def scene_caps_lean(scene) -> bool:
    """True when the scene runs the lean kernel instantiation (flat + conic surfaces only, OTB_CAPS_LENS)"""
    kinds = {r["kind"] for r in scene.flat.surfaces} if scene is not None else set()
    return not (kinds & {5, 6, 7, 8})
from . import _state
from . import color


class Raytracer(Group):

    N_EPS: float = 1e-11
    HURB_FACTOR: float = 2**0.5
    MAX_RAY_STORAGE_RAM: int = 150_000_000_000
    """maximum ray-storage bytes per GPU (the reference's host limit is 6 GB, raytracer.py:37; a B200 has 180 GB)"""
    upload_every_trace: bool = False
    """re-send scene and sampling tables host -> device on every trace even when unchanged (used by bench.py's
    end-to-end measurement, whose timed region must contain the host -> device copy of the step's inputs)"""
    deferred_status: bool = False
    """extension: trace() / trace_rays() return as soon as the kernels are queued instead of waiting for the device
    status word and the message counters (the one host synchronisation of a trace).  Both are collected by the next
    call that synchronises anyway — detector_image / detector_spectrum after their hit search, the next trace,
    source_image, focus_search — or by finish_trace(); a device-side error (hit-finding timeout, n < 1, ...) is
    raised there.  Default False: errors and messages appear inside trace() like in the reference."""
    use_specialised_kernels: bool = True
    arithmetic: str = "exact"
    """floating-point contract of the lens-surface step on the device.  "exact" (default): every + - * / sqrt rounds
    like the reference's numpy float64 operation, results are bit-identical to the reference on closed-form
    scenes.  "relaxed": fused multiply-adds, reciprocal-multiply division, rsqrt-based normalisation — 1.2x faster,
    each operation within 1-2 ulp, but NOT a drop-in at the 1e-9 level: the reference's own sphere intersection
    (B^2 - C) cancels 7 digits for sources tens of metres away, so any other rounding moves hit points by up to
    ~1e-9 relative, and float32-stored weights / polarisation flip by a float32 ulp (6e-8)."""
    """use a cached scene-specialised engine build when one exists (see Raytracer.compile)"""
    ITER_RAYS_STEP: int = 8_000_000
    """rays per iterative_render chunk (reference: 1e6, raytracer.py:40; larger chunks keep all 148 SMs busy)"""

    class INFOS(IntEnum):
        ABSORB_MISSING = 0
        TIR = 1
        ILL_COND = 2
        OUTLINE_INTERSECTION = 3
        HURB_NEG_DIR = 4

    def __init__(self, outline, n0: RefractionIndex = None, no_pol: bool = False, use_hurb: bool = False, **kwargs):
        self.outline = outline
        self.no_pol = no_pol
        self.use_hurb = use_hurb
        self.rays = RayStorage()
        self._msgs = np.array([])
        self._ignore_geometry_error = False
        self.geometry_error = False
        self._last_trace_snapshot = None
        self.fault_pos = np.array([])
        self._scene = None          # engine.SceneHandle of the last trace
        self._scene_key = None
        self._gen_cache = None
        self.seed = 0x0B200         # Philox key of the device generator; bump for independent bundles
        self._trace_count = 0
        super().__init__(None, n0, **kwargs)

    def __setattr__(self, key, val):
        if key == "outline":
            if not isinstance(val, (list, np.ndarray)):
                raise TypeError("outline needs to be a list or array.")
            o = np.asarray_chkfinite(val, dtype=np.float64)
            if o.shape[0] != 6 or o[0] >= o[1] or o[2] >= o[3] or o[4] >= o[5]:
                raise ValueError("Outline needs to be specified as [x1, x2, y1, y2, z1, z2] "
                                 "with x2 > x1, y2 > y1, z2 > z1.")
            val = o
        elif key in ("no_pol", "use_hurb") and not isinstance(val, bool):
            raise TypeError(f"{key} needs to be bool.")
        elif key == "arithmetic" and val not in ("exact", "relaxed"):
            raise ValueError("arithmetic needs to be 'exact' or 'relaxed'.")
        object.__setattr__(self, key, val)

    @property
    def extent(self):
        return tuple(self.outline)

    @property
    def pos(self):
        return np.mean(self.outline[:2]), np.mean(self.outline[2:4]), self.outline[4]

    def clear(self) -> None:
        super().clear()
        self.rays = RayStorage()

    # -- change detection (raytracer.py:141-179): structural hash of the flattened scene -------------
    def _geometry_state(self):
        """plain-value state of everything the flattened scene depends on.  The deep walk (~0.1 ms) is reused while
        no attribute of any scene object was assigned (scene epoch, _state.py); positions are re-read every time
        so that in-place edits of a `pos` array are still noticed."""
        raw = self._elements          # unsorted: the sort by z (a `pos` evaluation per element) is not needed for a key
        quick = (_state.EPOCH[0], len(raw), tuple(self.outline), self.no_pol, self.use_hurb, self.HURB_FACTOR,
                 self.arithmetic,
                 id(self.n0), tuple((id(el), el._front.pos.tobytes()) for el in raw),
                 tuple((id(rs), id(rs.or_func)) for rs in self.ray_sources if rs.orientation == "Function"))
        cache = self.__dict__.get("_geom_cache")
        if cache is not None and cache[0] == quick:
            return cache[1]
        els = tuple((type(el).__name__, state_of(el._front), state_of(el._back), el._d1, el._d2,
                     state_of(getattr(el, "n", None)), state_of(getattr(el, "n2", None)),
                     state_of(getattr(el, "spectrum", None)), getattr(el, "D", None))
                    for el in self.elements if isinstance(el, (Lens, Filter, Aperture)))
        key = (els, tuple(self.outline), state_of(self.n0), self.no_pol, self.use_hurb, self.HURB_FACTOR, self.arithmetic,
               quick[-1])
        object.__setattr__(self, "_geom_cache", (quick, key))
        return key

    def tracing_snapshot(self, scene_key=None):
        src = [(id(rs), tuple(rs.pos), rs.power, rs.divergence, rs.orientation, rs.polarization, rs.div_angle,
                tuple(rs.s), tuple(rs.conv_pos), id(rs.spectrum)) for rs in self.ray_sources]
        return dict(scene=scene_key if scene_key is not None else self._geometry_state(), sources=src,
                    rays=self.rays.crepr(),
                    settings=(self.no_pol, self.use_hurb, self.HURB_FACTOR))

    def check_if_rays_are_current(self) -> bool:
        if self._last_trace_snapshot is None:
            return False
        return self._last_trace_snapshot == self.tracing_snapshot()

    # -- messages (raytracer.py:192-244) ----------------------------------------------------------------
    def _surface_names(self):
        names = {}
        for type_, els in zip(["Lens", "Aperture", "Filter"], [self.lenses, self.apertures, self.filters]):
            for i, el in enumerate(els):
                if not el.has_back() or getattr(el, "is_ideal", False):
                    names[f"surface of {type_} {el.abbr}{i}"] = el.pos[2]
                else:
                    names[f"front surface of {type_} {el.abbr}{i}"] = el.front.pos[2]
                    names[f"back surface of {type_} {el.abbr}{i}"] = el.back.pos[2]
        return ["RaySource"] + sorted(names, key=lambda k: names[k]) + ["Outline"]

    def _show_messages(self, N) -> None:
        if not global_options.show_warnings or not self._msgs.size or not self._msgs.any():
            return
        names = self._surface_names()
        text = {self.INFOS.TIR: "with total inner reflection at surface {s} ({n}), treating as absorbed.",
                self.INFOS.ABSORB_MISSING: "missing lens surface {s} ({n}), set to absorbed",
                self.INFOS.ILL_COND: "are ill-conditioned for numerical hit finding at surface {s} ({n}). "
                                     "Where and whether they intersect might be wrong.",
                self.INFOS.OUTLINE_INTERSECTION: "hitting outline after surface {s} ({n}), set to absorbed.",
                self.INFOS.HURB_NEG_DIR: "have negative z-direction after ray bending at surface {s} ({n}), "
                                         "set to absorbed."}
        for t in range(self._msgs.shape[0]):
            for s in range(self._msgs.shape[1]):
                if count := int(self._msgs[t, s]):
                    nm = names[s] if s < len(names) else "?"
                    warning(f"{count} rays ({100*count/N:.3g}% of all rays) " + text[t].format(s=s, n=nm))

    # -- geometry checks (raytracer.py:510-664) ---------------------------------------------------------------
    @staticmethod
    def check_collision(front, back, res: int = 100):
        """Raytracer.check_collision (raytracer.py:581-664): does `front` reach behind `back` somewhere both are
        defined?  Returns (collision flag, x, y, z arrays of the colliding samples).  Point / line against surface:
        the surface height at the point(s); surface against surface: a res x res grid over the overlap of the two
        xy extents, compared where both masks hold.  Scene-setup sampling on the host (1e4 samples per pair)."""
        from .surfaces import Surface, Point, Line
        none = (False, np.array([]), np.array([]), np.array([]))
        if not (isinstance(front, Surface) or isinstance(back, Surface)):
            raise TypeError("At least one object needs to be a Surface for collision detection")
        for kind in (Point, Line):
            if isinstance(front, kind) or isinstance(back, kind):
                first = isinstance(front, kind)           # the point / line is the object in front
                obj, surf = (front, back) if first else (back, front)
                if kind is Point:
                    x, y = np.array([obj.pos[0]]), np.array([obj.pos[1]])
                else:
                    t = np.linspace(-obj.r, obj.r, 10*res)
                    x, y = obj.pos[0] + np.cos(obj.angle)*t, obj.pos[1] + np.sin(obj.angle)*t
                z = np.asarray(surf.values(x, y), dtype=np.float64)
                bad = ((z < obj.pos[2]) if first else (z > obj.pos[2])) & surf.mask(x, y)
                k = np.nonzero(bad)[0]
                return bool(bad.any()), x[k], y[k], z[k]
        ef, eb = front.extent, back.extent
        if ef[5] < eb[4]:                                  # z extents apart
            return none
        x0, x1, y0, y1 = max(ef[0], eb[0]), min(ef[1], eb[1]), max(ef[2], eb[2]), min(ef[3], eb[3])
        if x0 > x1 or y0 > y1:                             # no common area in the xy projection
            return none
        Y, X = np.mgrid[y0:y1:res*1j, x0:x1:res*1j]
        x, y = X.flatten(), Y.flatten()
        both = front.mask(x, y) & back.mask(x, y)
        x, y = x[both], y[both]
        zf = np.asarray(front.values(x, y), dtype=np.float64)
        zb = np.asarray(back.values(x, y), dtype=np.float64)
        k = np.nonzero(zf > zb)[0]
        return bool(k.shape[0]), x[k], y[k], zf[k]

    def _tracing_elements(self):
        """z-ordered Lens / Filter / Aperture elements plus the end absorber at the outline (raytracer.py:492-508)"""
        o = self.outline
        end = Aperture(RectangularSurface(dim=[o[1] - o[0], o[3] - o[2]]), pos=[(o[1] + o[0])/2, (o[2] + o[3])/2, o[5]])
        return [el for el in self.elements if isinstance(el, (Lens, Filter, Aperture))] + [end]

    def _geometry_checks(self) -> None:
        """Raytracer.__geometry_checks (raytracer.py:510-578): elements and sources inside the outline, surfaces in
        sequence without collisions (front | back of every element, every element | the next one, sources | first
        element), HURB only on ring / slit apertures.  Sets geometry_error and fault_pos."""
        o = self.outline + self.N_EPS*np.array([-1, 1, -1, 1, -1, 1])

        def inside(e):
            return o[0] <= e[0] and e[1] <= o[1] and o[2] <= e[2] and e[3] <= o[3] and o[4] <= e[4] and e[5] <= o[5]

        def fail(msg=None):
            if msg:
                warning(msg)
            self.geometry_error = True

        if not self.ray_sources:
            return fail("RaySource Missing.")
        els = self._tracing_elements()
        hit = None
        for i, el in enumerate(els):
            if not inside(el.extent):
                return fail(f"Element{i} {el} with extent {el.extent} outside outline {self.outline}.")
            pairs = []
            if i + 1 < len(els):
                pairs.append((el.front, els[i + 1].front))
            if el.has_back():
                pairs.append((el.front, el.back))
                pairs.append((el.back, els[i + 1].front))
            for a, b in pairs:
                c = self.check_collision(a, b)
                if c[0]:
                    hit = c
                    break
            if self.use_hurb and i < len(els) - 1 and isinstance(el, Aperture) \
                    and not isinstance(el.front, (RingSurface, SlitSurface)):
                return fail(f"Ray bending for surface type {type(el.front).__name__} not implemented.")
            if hit:
                break
        if not hit:
            for rs in self.ray_sources:
                if not inside(rs.extent):
                    return fail(f"RaySource {rs} with extent {rs.extent} outside outline {self.outline}.")
                if rs.pos[2] >= els[0].extent[4]:
                    c = self.check_collision(rs.surface, els[0].front)
                    if c[0]:
                        hit = c
                        break
        if hit:
            _, xc, yc, zc = hit
            fail(f"Detected collision between two Surfaces at {xc[0], yc[0], zc[0]}"
                 f" and at least {xc.shape[0]} other positions.")
            self.fault_pos = np.column_stack((xc, yc, zc))
            return
        self.geometry_error = False

    def _pretrace_check(self, N) -> bool:
        if not isinstance(N, int) or isinstance(N, bool):
            raise TypeError(f"N needs to be of type int, but is {type(N)}.")
        if N < 1:
            raise ValueError(f"Ray number N needs to be at least 1, but is {N}.")
        # the checks depend on the geometry and the sources only: repeated traces of an unchanged scene reuse the
        # verdict (the warnings were shown when it was computed)
        ck = (self._geometry_state(), tuple((id(rs), rs._front.pos.tobytes()) for rs in self.ray_sources))
        cache = self.__dict__.get("_check_cache")
        if cache is not None and cache[0] == ck and not cache[1]:
            self.geometry_error = False
        else:
            self._geometry_checks()
            object.__setattr__(self, "_check_cache", (ck, bool(self.geometry_error)))
        if self.geometry_error and not self._ignore_geometry_error:
            warning("ABORTED TRACING")
            return True
        return False

    # -- scene / generator upload --------------------------------------------------------------------
    def _scene_handle(self, specialised="cached"):
        key = self._geometry_state()
        if self._scene is not None and self._scene_key == key and not (specialised is True and not self._scene.specialised):
            if self.upload_every_trace:     # re-send the (unchanged) scene tables host -> device
                self._scene.reupload()
            return self._scene              # geometry unchanged: no re-flattening
        flat = flatten_raytracer(self)
        if self._scene is not None:
            self._scene.close()
        self._scene = engine.SceneHandle(flat, specialised if self.use_specialised_kernels else False)
        self._scene_key = key
        return self._scene

    def compile(self) -> bool:
        """Extension of the reference API: build (once, ~40 s of nvcc, cached in-tree) and select trace kernels
        specialised for the current geometry — scene baked in as a device constant, step loop unrolled
        (optrace_b200/specialise.py).  Results are bit-identical to the generic kernels; worth it for scenes
        traced many times (iterative renders of billions of rays, sweeps, serving).  Later traces of the same
        geometry pick the cached build up automatically."""
        engine.ensure_init()
        return self._scene_handle(specialised=True).specialised

    def _generator_tables(self):
        """per-source generator records and the shared table buffer on the device (cached per source set)"""
        torch = engine._torch()
        key = tuple((id(rs), id(rs.spectrum), rs.power, rs.divergence, rs.div_angle, rs.orientation, rs.polarization,
                     tuple(rs.pos), tuple(rs.s), tuple(rs.conv_pos), rs.div_2d, rs.pol_angle, rs.div_axis_angle)
                    for rs in self.ray_sources) + (tuple(global_options.wavelength_range),)
        if self._gen_cache is not None and self._gen_cache[0] == key:
            if self.upload_every_trace:
                self._gen_cache[2].copy_(self._gen_cache[3], non_blocking=True)
            return self._gen_cache[1], self._gen_cache[2]
        recs, chunks, off = [], [], 0

        def guided(x, F, kind):
            """x, F + guide table G for the device lookups (otb_gen.cu): bracket at n equidistant CDF levels"""
            F = np.asarray(F, dtype=np.float64)
            n = F.shape[0]
            if kind == "linear":
                edges = F[0] + (F[-1] - F[0])*np.arange(n)/n
                G = np.clip(np.searchsorted(F, edges, side="right") - 1, 0, max(n - 2, 0))
            else:
                G = np.clip(np.searchsorted(F, F[-1]*np.arange(n)/n, side="left"), 0, n - 1)
            return (x, F, G.astype(np.float64))

        def put(arrs):
            nonlocal off
            a = np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in arrs])
            chunks.append(a)
            o = off
            off += a.shape[0]
            return o

        srgb_off = -1
        for rs in self.ray_sources:
            r = rs._generator_record()
            t = r["tables"]
            wl_kind = "next" if r["wl"]["mode"] == 2 else "linear"
            pol_kind = "next" if r["polarization"] == 2 else "linear"
            r["wl_tab_off"], r["wl_tab_n"] = (put(guided(*r["wl"]["tab"], wl_kind)), len(r["wl"]["tab"][0])) \
                if r["wl"]["tab"] is not None else (0, 0)
            r["div_tab_off"], r["div_tab_n"] = (put(guided(*t["div"], "linear")), len(t["div"][0])) if "div" in t else (0, 0)
            r["pol_tab_off"], r["pol_tab_n"] = (put(guided(*t["pol"], pol_kind)), len(t["pol"][0])) if "pol" in t else (0, 0)
            r["pix_cdf_off"], r["pix_cdf_n"] = (put(guided(t["pix_idx"], t["pix_cdf"], "next")), len(t["pix_idx"])) \
                if "pix_idx" in t else (0, 0)
            r["pix_rgb_off"] = put((t["pix_rgb"],)) if "pix_rgb" in t else 0
            if r["shape"] == 5:
                if srgb_off < 0:
                    wl5, Fr, Fg, Fb = color.srgb_primary_cdfs()
                    srgb_off = put(sum((guided(wl5, F, "linear") for F in (Fr, Fg, Fb)), ()))
                r["srgb_off"] = srgb_off
            else:
                r["srgb_off"] = 0
            recs.append(r)
        aux = np.concatenate(chunks) if chunks else np.zeros(1)
        aux_h = torch.from_numpy(np.ascontiguousarray(aux)).pin_memory()
        aux_d = aux_h.to(engine.device(), non_blocking=True)
        self._gen_cache = (key, recs, aux_d, aux_h)
        return recs, aux_d

    coherent_bundles: bool = True
    """order of the rays inside a generated bundle.  The reference shuffles every stratified sample (random.py:41-45);
    with this flag (default) the strata of ONE random variable — the direction inside the divergence cone when the
    source has one, else the position on the source area — are handed out in blocks of 32 neighbouring cells, so the
    32 rays of a warp meet stops and lens edges together and warps whose rays are all absorbed skip the surface
    arithmetic.  Same cells, same distributions, same images; only the order of the rays within a source block is
    less random (a slice rays[:k] is a set of 32-ray clusters instead of a uniform subsample).  False = full shuffle."""

    def _source_records(self, scene, N_list, blocks):
        """OtbSource array of the local blocks (source, global first ray id, count) + the table buffer on the device"""
        recs, aux_d = self._generator_tables()          # (re-)upload of the sampling tables on the current stream
        sl = blocks
        fids = scene.flat.source_func_ids
        coherent = int(bool(self.coherent_bundles) and scene_caps_lean(scene))
        # the records are plain host structures: rebuilt only when something they are made of changed (a repeated
        # trace of the same scene re-sends the device tables above, not this loop of ~40 field assignments per source)
        key = (self._gen_cache[0], tuple(int(v) for v in N_list), tuple(sl), coherent, tuple(sorted(fids.items())))
        cached = self.__dict__.get("_source_rec_cache")
        if cached is not None and cached[0] == key:
            return cached[1], len(sl), aux_d
        arr = (_cabi.OtbSource*max(len(sl), 1))()
        start = 0
        for k, (i, gid0, cnt) in enumerate(sl):
            r, S = recs[i], arr[k]
            S.gid_start = gid0
            S.shape, S.orientation, S.divergence = r["shape"], r["orientation"], r["divergence"]
            S.polarization, S.wl_mode, S.div_2d = r["polarization"], r["wl"]["mode"], r["div_2d"]
            S.img_w, S.img_h = r["img_w"], r["img_h"]
            S.n_rays, S.ray_start = cnt, start
            start += cnt
            S.power = r["power"]
            # ray_source.py:219-220 / ray_storage.py:160-163: float32(power/N) with the per-slice power share
            S.weight = float(np.float32(r["power"]/N_list[i])) if N_list[i] else 0.0
            S.pos[:], S.geom[:], S.extent[:] = r["pos"], r["geom"], r["extent"]
            S.s[:], S.conv_pos[:] = r["s"], r["conv_pos"]
            S.div_sin, S.div_angle, S.div_axis, S.pol_angle = r["div_sin"], r["div_angle"], r["div_axis"], r["pol_angle"]
            S.wl[:] = [float(v) for v in r["wl"]["wl"]]
            for f in ("wl_tab_off", "wl_tab_n", "div_tab_off", "div_tab_n", "pol_tab_off", "pol_tab_n",
                      "pix_cdf_off", "pix_cdf_n", "pix_rgb_off", "srgb_off"):
                setattr(S, f, int(r[f]))
            S.or_func_id = fids.get(id(self.ray_sources[i]), -1)
            # measured: the numeric-surface kernels (CAPS_FULL) are bound by instruction fetch and lose with warps that
            # run different iteration counts side by side (cosine_surfaces 11.6 -> 14.7 ms): coherent order only for
            # scenes of flat and conic surfaces
            S.coherent = coherent
        object.__setattr__(self, "_source_rec_cache", (key, arr))
        return arr, len(sl), aux_d

    def _generate(self, N_list, begin, end: int, seed: int, scene=None):
        """otb_generate_rays for the local shard [begin, end) of the global ray range: the stand-alone generator
        (tests, tools, injected-bundle workflows).  trace() and iterative_render() do not call it: they hand the
        source records to the trace kernels, which draw the rays themselves (engine.DeviceRays.generated)."""
        torch = engine._torch()
        engine.ensure_init()
        scene = scene or self._scene_handle()
        lib = scene.lib
        # `begin` may be the list of local blocks (balanced sharding); else the contiguous global range [begin, end)
        blocks = begin if isinstance(begin, list) else dist.contiguous_blocks(N_list, begin, end)
        n = sum(c for _, _, c in blocks)
        begin = self._ray_offset(N_list, blocks, begin)
        arr, ns, aux_d = self._source_records(scene, N_list, blocks)
        d = engine.device()
        p0 = torch.empty(3*n, dtype=torch.float64, device=d)
        s0 = torch.empty(3*n, dtype=torch.float64, device=d)
        pol0 = None if self.no_pol else torch.empty(3*n, dtype=torch.float32, device=d)
        w0 = torch.empty(n, dtype=torch.float32, device=d)
        wl = torch.empty(n, dtype=torch.float32, device=d)
        status = torch.zeros(1, dtype=torch.int32, device=d)
        _cabi.check(lib.otb_generate_rays(arr, ns, engine.dptr(aux_d), n, seed, begin, int(self.no_pol),
                                          engine.dptr(p0), engine.dptr(s0), engine.dptr(pol0), engine.dptr(w0),
                                          engine.dptr(wl), engine.dptr(status), engine.stream_ptr()), lib)
        rays = engine.DeviceRays(n, p0, s0, pol0, w0, wl, None, seed, begin)
        rays.gen_status = status        # checked when the trace result is synchronised anyway
        return rays

    fused_generation: bool = False
    """draw the rays inside the trace kernels (OtbRays.gen_h) instead of with the generator kernel.  A fused bundle
    never exists in HBM (68 B per ray less memory and traffic, identical rays bit for bit), but the trace kernels run
    at 16 warps per SM and the generator's integer work does not hide behind their fp64 chains: measured on the
    double-Gauss workload 4.53 ms fused against 3.67 + 0.66 ms separate.  Worth it when device memory is the limit."""

    def _generated(self, scene, N_list, begin, end: int, seed: int):
        """the bundle of a trace: generated by the generator kernel (default) or described for the fused path.
        `begin`: list of local blocks (dist.shard_sources) or the start of the contiguous range [begin, end)"""
        if not self.fused_generation:
            return self._generate(N_list, begin, end, seed, scene)
        blocks = begin if isinstance(begin, list) else dist.contiguous_blocks(N_list, begin, end)
        arr, ns, aux_d = self._source_records(scene, N_list, blocks)
        return engine.DeviceRays.generated(sum(c for _, _, c in blocks), arr, ns, aux_d, seed,
                                           self._ray_offset(N_list, blocks, begin))

    @staticmethod
    def _ray_offset(N_list, blocks, begin) -> int:
        """counter offset of the per-ray draws made inside the trace kernels (HURB deviates: Philox counter =
        offset + local ray index): the global id of the first ray for a contiguous range; for the balanced shards of a
        job (a slice of every source) the number of rays the lower ranks hold, so that no two rays share a counter"""
        if not isinstance(begin, list):
            return int(begin)
        return dist.rank()*sum(int(n)//dist.world() for n in N_list)

    # -- trace -----------------------------------------------------------------------------------------
    def trace(self, N: int) -> None:
        """Raytracer.trace (raytracer.py:262-415): N rays, generated and traced on the GPU(s)."""
        self.finish_trace()
        if self._pretrace_check(N):
            return
        engine.ensure_init()
        scene = self._scene_handle()
        nt = scene.nt
        N_list = dist.shared_split(N, [rs.power for rs in self.ray_sources], engine.device())
        blocks = dist.shard_sources(N_list)          # this rank's slice of every source (balanced work per rank)
        n_local = sum(c for _, _, c in blocks)
        if RayStorage.storage_size(n_local, nt, self.no_pol) > self.MAX_RAY_STORAGE_RAM:
            raise RuntimeError(f"More than {self.MAX_RAY_STORAGE_RAM*1e-9:.1f} GB RAM requested. Either decrease"
                               " the number of rays, surfaces or do an iterative render. If your system can handle"
                               " more RAM usage, increase the Raytracer.MAX_RAY_STORAGE_RAM parameter.")
        if np.any(N_list == 0):
            warning("There are RaySources that have no rays assigned. "
                    "Change the power ratio or raise the overall ray number")
        self._trace_count += 1
        seed = (int(self.seed) << 20) + self._trace_count
        rays = self._generated(scene, N_list, blocks, 0, seed)
        self._run_trace(scene, rays, N_list, N, blocks)

    def trace_rays(self, p, s, pol, w, wl, hurb_z=None, N_list=None, sharded: bool = False) -> None:
        """Extension of the reference API: trace a pre-generated bundle (the arrays RaySource.create_rays
        returns: p, s float64 (N,3); pol (N,3) or None with no_pol; w, wl (N)).  `hurb_z` (n_hurb, 2, N)
        optionally injects the standard normal deviates of the HURB bending.  Used for parity runs on
        identical bundles.  `sharded`: under torchrun every rank passes the SAME global bundle
        and traces its contiguous share of it (dist.shard_range), like trace(N) does with generated rays."""
        N = int(p.shape[0])
        self.finish_trace()
        if self._pretrace_check(N):
            return
        engine.ensure_init()
        scene = self._scene_handle()
        if self.no_pol:
            pol = None
        elif pol is None:
            raise ValueError("pol is required unless no_pol is set.")
        begin, end = dist.shard_range(N) if sharded else (0, N)
        if (begin, end) != (0, N):
            p, s, w, wl = p[begin:end], s[begin:end], w[begin:end], wl[begin:end]
            pol = None if pol is None else pol[begin:end]
            hurb_z = None if hurb_z is None else np.ascontiguousarray(np.asarray(hurb_z)[:, :, begin:end])
        rays = engine.DeviceRays.from_host(p, s, pol, w, wl, hurb_z, seed=int(self.seed))
        N_list = np.array([N]) if N_list is None else np.asarray(N_list, dtype=int)
        self._run_trace(scene, rays, N_list, N, begin)

    def _run_trace(self, scene, rays, N_list, N_global, begin):
        # drop the previous ray storage first: its device blocks go back to the caching allocator and the new
        # store (same size on a repeated trace) reuses them instead of a fresh cudaMalloc of many GB
        # When nobody but this tracer can still see the previous RayStorage, its device planes are overwritten in
        # place by the new trace (same ray count and section count): returning 8 GB to the caching allocator and asking
        # for them again makes every change of the set of live tensors (an image kept, a pipeline drained) a candidate
        # for a synchronous multi-GB cudaMalloc (8-400 ms measured).  A RayStorage somebody else holds keeps its planes.
        old = self.rays
        recycled = None
        if sys.getrefcount(old) <= 3:                       # self.rays, `old`, the argument of getrefcount
            dev_store = old.__dict__.get("_dev")
            if dev_store is not None and sys.getrefcount(dev_store) <= 3 and \
                    (dev_store.N, dev_store.nt, dev_store.no_pol) == (rays.N, scene.nt, scene.flat.no_pol):
                recycled = dev_store
        del old
        self.rays = RayStorage()
        store, msgs, status = engine.trace_store(scene, rays, store=recycled, sync=False)
        recycled = None
        gen_status = getattr(rays, "gen_status", None)
        if gen_status is not None:
            status = status | gen_status
        # the one collective and the one host synchronisation of a trace: message counters summed, status words
        # OR-ed over all ranks, so that every rank raises the same exception
        self._pending_trace = (dist.reduce_msgs_status_begin(msgs, status), N_global)
        if not self.deferred_status:
            self.finish_trace()
        self.rays = RayStorage()
        self.rays._attach(store, self.ray_sources, N_list, self.no_pol, N_global, begin)
        self._last_trace_snapshot = self.tracing_snapshot(self._scene_key)     # flattened once per trace

    def finish_trace(self) -> None:
        """collect the status word and the message counters of the last trace (see `deferred_status`); a no-op when
        that has happened already"""
        pend = self.__dict__.get("_pending_trace")
        if pend is None:
            return
        self._pending_trace = None
        handle, N_global = pend
        self._msgs, st = dist.reduce_msgs_status_end(handle)
        engine.raise_status(st)
        self._show_messages(N_global)

    # -- detector ----------------------------------------------------------------------------------------
    def _check_detector_call(self, detector_index, source_index):
        if not self.detectors:
            raise RuntimeError("Detector Missing")
        if not self.rays.N_global:
            raise RuntimeError("No rays traced.")
        if source_index is not None and (source_index > len(self.ray_sources) - 1 or source_index < 0):
            raise IndexError("Invalid source_index.")
        if detector_index > len(self.detectors) - 1 or detector_index < 0:
            raise IndexError("Invalid detector_index.")
        if not self.check_if_rays_are_current():
            raise RuntimeError("Tracing geometry/properties changed. Please retrace first.")

    def _hit_detector(self, detector_index=0, source_index=None, extent=None, projection_method="Equidistant"):
        """Raytracer._hit_detector (raytracer.py:881-1051) on the device.  Returns device tensors
        (hx, hy, hw, wl) over the selected local ray range plus (extent_out, projection, ill_count)."""
        self._check_detector_call(detector_index, source_index)
        if extent is not None and not isinstance(extent, (list, np.ndarray)):
            raise ValueError(f"Invalid extent '{extent}'.")
        dsurf = self.detectors[detector_index].surface
        rec = detector_record(dsurf, projection_method, extent)
        b, e = self.rays._local_range(source_index)
        lib = self._scene.lib
        hx, hy, hw, rng, ill, status, meta = engine.detector_hits(lib, self.rays._dev, rec, b, e)
        projection = projection_method if rec["projection"] else None
        # hit range (MIN / MAX), ill-conditioned count (SUM) and status (OR) of all ranks: one all-gather of the
        # 48-byte record and the one host synchronisation of this call, reduced on the host
        r, ill_count, st = engine.read_det_meta(meta)
        self.finish_trace()            # deferred status of the trace: the device has just been synchronised
        if extent is not None:
            extent_out = np.asarray_chkfinite(np.array(extent, dtype=np.float64))
        else:
            extent_out = self.detectors[detector_index].pos[:2].repeat(2)
            if r[0] <= r[1]:
                extent_out = r.copy()
        engine.raise_status(st)
        return hx, hy, hw, self.rays._dev.wl[b:e], extent_out, projection, ill_count

    def detector_image(self, detector_index: int = 0, source_index: int = None, extent=None, limit: float = None,
                       projection_method: str = "Equidistant", **kwargs) -> RenderImage:
        """Raytracer.detector_image (raytracer.py:1053-1098)"""
        dont_filter = bool(kwargs.get("_dont_filter", False))
        if limit is not None and extent is not None and not dont_filter:
            warning("Using the limit parameter in combination with a user defined extent"
                    " will produce an incorrect detector image, as the rays outside the extent"
                    " are not included in the convolution calculation.")
        hx, hy, hw, wl, extent_out, projection, ill_count = \
            self._hit_detector(detector_index, source_index, extent, projection_method)
        det = self.detectors[detector_index]
        pname = f": {det.desc}" if det.desc != "" else ""
        desc = f"{Detector.abbr}{detector_index}{pname} at z = {det.pos[2]:.5g} mm"
        if source_index is not None:
            desc = f"Rays from RS{source_index} at " + desc
        img = RenderImage(long_desc=desc, extent=extent_out, projection=projection)
        img._limit = limit                      # enlarges the extent by 2.7 limit (render_image.py:250-252)
        img._fix_extent()
        Nx, Ny = img._grid()
        data, cnt = engine.render_xyzw(self._scene.lib, hx, hy, hw, wl, img.extent, Nx, Ny)
        img._data_dev, img._counts_dev, img._lib = data, cnt, self._scene.lib
        img._ready, img._pack = dist.allreduce_image_async(self._scene.lib, data)       # side stream, occupied tiles only
        img._counts_local = dist.world() > 1      # the count channel (parity checks) is reduced on first access
        if limit is not None and not dont_filter:
            img._apply_rayleigh_filter()        # resolution filter on the reduced image (render_image.py:420-421)
        if ill_count:
            warning(f"{ill_count} rays ({100*ill_count/self.rays.N_global:.3g}% of all rays) were ill-conditioned for "
                    f"numerical hit finding at detector {detector_index}. Where and whether they intersect might be wrong.")
        return img

    def detector_spectrum(self, detector_index: int = 0, source_index: int = None, extent=None, **kwargs) -> LightSpectrum:
        """Raytracer.detector_spectrum (raytracer.py:1100-1132); weighted histogram of the hit wavelengths on the
        device (otb_spectrum_stats / otb_spectrum_hist), counts, range and bin sums reduced over all GPUs"""
        hx, hy, hw, wl, _, _, ill_count = self._hit_detector(detector_index, source_index, extent)
        det = self.detectors[detector_index]
        pname = f": {det.desc}" if det.desc != "" else ""
        desc = f"{Detector.abbr}{detector_index}{pname} at z = {det.pos[2]:.5g} mm"
        desc = (f"Spectrum of RS{source_index} at " if source_index is not None else "Spectrum at ") + desc
        spec = LightSpectrum._render_device(self._scene.lib, wl, hw, True, long_desc=desc, **kwargs)
        if ill_count:
            warning(f"{ill_count} rays ({100*ill_count/self.rays.N_global:.3g}% of all rays) were ill-conditioned for "
                    f"numerical hit finding at detector {detector_index}. Where and whether they intersect might be wrong.")
        return spec

    def source_image(self, source_index: int = 0, limit: float = None, **kwargs) -> RenderImage:
        """Raytracer.source_image (raytracer.py:1331-1352)"""
        self.finish_trace()
        if not self.ray_sources:
            raise RuntimeError("Ray Sources Missing.")
        if not self.rays.N_global:
            raise RuntimeError("No rays traced.")
        if source_index > len(self.ray_sources) - 1 or source_index < 0:
            raise IndexError("Invalid source_index.")
        if not self.check_if_rays_are_current():
            raise RuntimeError("Tracing geometry/properties changed. Please retrace first.")
        rs = self.ray_sources[source_index]
        b, e = self.rays._local_range(source_index)
        st = self.rays._dev
        N, nt = st.N, st.nt
        x = st.p[b:e]
        y = st.p[N*nt + b:N*nt + e]
        img = RenderImage(long_desc=f"{RaySource.abbr}{source_index} at z = {rs.pos[2]:.5g} mm",
                          extent=np.array(rs.extent[:4], dtype=np.float64), projection=None)
        img._limit = limit
        img._fix_extent()
        Nx, Ny = img._grid()
        data, cnt = engine.render_xyzw(self._scene.lib, x, y, st.w[b:e], st.wl[b:e], img.extent, Nx, Ny)
        img._data_dev, img._counts_dev, img._lib = data, cnt, self._scene.lib
        img._ready, img._pack = dist.allreduce_image_async(self._scene.lib, data)
        img._counts_local = dist.world() > 1
        if limit is not None and not kwargs.get("_dont_filter", False):
            img._apply_rayleigh_filter()
        return img

    def source_spectrum(self, source_index: int = 0, **kwargs) -> LightSpectrum:
        """Raytracer.source_spectrum (raytracer.py:1311-1329), binned on the device like detector_spectrum"""
        self.finish_trace()
        if not self.ray_sources:
            raise RuntimeError("Ray Sources Missing.")
        if not self.rays.N_global:
            raise RuntimeError("No rays traced.")
        if source_index > len(self.ray_sources) - 1 or source_index < 0:
            raise IndexError("Invalid source_index.")
        if not self.check_if_rays_are_current():
            raise RuntimeError("Tracing geometry/properties changed. Please retrace first.")
        b, e = self.rays._local_range(source_index)
        st = self.rays._dev
        rs = self.ray_sources[source_index]
        pname = f": {rs.desc}" if rs.desc != "" else ""
        return LightSpectrum._render_device(self._scene.lib, st.wl[b:e], st.w[b:e], False,
                                            long_desc=f"Spectrum of {RaySource.abbr}{source_index}{pname} at z = "
                                                      f"{rs.pos[2]:.5g} mm", **kwargs)

    # -- focus search (raytracer.py:1354-1640) ---------------------------------------------------------------
    focus_search_methods = ["RMS Spot Size", "Irradiance Variance", "Image Sharpness", "Image Center Sharpness"]

    @staticmethod
    def _focus_image_cost(mode: str, ext: np.ndarray, Im: np.ndarray) -> float:
        """cost of one rendered N_px x N_px image (raytracer.py:1394-1420): O(N_px^2) host arithmetic"""
        N_px = Im.shape[0]
        if mode in ("Image Sharpness", "Image Center Sharpness"):
            if mode == "Image Center Sharpness":
                Y, X = np.mgrid[-1:1:N_px*1j, -1:1:N_px*1j]
                R = np.sqrt(X**2 + Y**2)
                win = np.where(R > 1, 0, 1 + np.cos(R*np.pi))
                Im0 = Im*win
                if (Im0s := Im0.sum()):
                    Im0 *= 1/Im0s
            else:
                Im0 = Im
            rsm = ((Im0[1:] - Im0[:-1])**2).sum() + ((Im0[:, 1:] - Im0[:, :-1])**2).sum()
            return -rsm
        Im = Im[Im > 0]
        Ap = (ext[1] - ext[0])*(ext[3] - ext[2])/N_px**2
        return -np.log(Im.var()/Ap**2)

    def _focus_cost(self, L, z_pos: float, mode: str) -> float:
        """Raytracer.__focus_search_cost_function (raytracer.py:1354-1420); per-ray sums and binning on the device"""
        if mode == "RMS Spot Size":
            sw, sw2, swx, swy = L.moments(0, [z_pos])
            cx, cy, _, _ = L.moments(1, [z_pos, swx/sw, swy/sw])
            fact = sw - sw2/sw                      # np.cov(..., aweights=w), ddof = 1
            return float(np.sqrt(cx/fact + cy/fact))
        N_px = 100*int(1 + np.sqrt(L.n_use)/1500)
        N_px = N_px if N_px % 2 else N_px + 1
        ext, Im = L.image(z_pos, N_px)
        return float(self._focus_image_cost(mode, ext, Im))

    def focus_search(self, method: str, z_start: float, source_index: int = None, return_cost: bool = False):
        """Raytracer.focus_search (raytracer.py:1449-1640).  Bounds, sampling of the cost function and the scipy
        optimisers are the reference's; every pass over the rays (section selection, weighted moments, cost images)
        runs on the device (otb_focus_prepare / otb_focus_moments / otb_focus_image)."""
        import scipy.optimize
        self.finish_trace()
        if not (self.outline[4] <= z_start <= self.outline[5]):
            raise ValueError(f"Starting position z_start={z_start} outside raytracer z-outline range {self.outline[4:]}.")
        if method not in self.focus_search_methods:
            raise ValueError(f"Invalid method '{method}', should be one of {self.focus_search_methods}.")
        if not self.rays.N_global:
            raise RuntimeError("No rays traced.")
        if source_index is not None and source_index < 0:
            raise IndexError(f"source_index needs to be >= 0, but is {source_index}")
        if (source_index is not None and source_index > len(self.rays.N_list)) or len(self.rays.N_list) == 0:
            raise IndexError(f"source_index={source_index} larger than number of simulated sources "
                             f"({len(self.rays.N_list)}. Either the source was not added or the new geometry was "
                             "not traced.")
        if not self.check_if_rays_are_current():
            raise RuntimeError("Tracing geometry/properties changed or last trace had errors. Please retrace first.")

        # search bounds: between the surfaces around z_start (raytracer.py:1505-1518)
        b0 = self.N_EPS + np.max([rs.extent[5] for rs in self.ray_sources])
        b1 = self.outline[5] - self.N_EPS
        for surf in self.tracing_surfaces:
            if surf.z_max > z_start:
                b1 = surf.z_min
                break
            b0 = surf.z_max
        bounds = [b0, b1]
        Nt = 320

        b, e = self.rays._local_range(source_index)
        L = engine.FocusLines(self._scene.lib, self.rays._dev, b, e, bounds[0] + self.N_EPS)
        N_use = L.n_use
        if N_use < 1000:
            warning(f"WARNING: Less than 1000 rays for focus_search ({N_use}).")
        if N_use <= 1:
            return scipy.optimize.OptimizeResult(), dict(pos=[np.nan, np.nan, np.nan], bounds=bounds, z=np.full(Nt, np.nan),
                                                         cost=np.full(Nt, np.nan), N=N_use)

        r = vals = None
        if return_cost or method in ("Image Sharpness", "Image Center Sharpness"):
            # random.stratified_interval_sampling(bounds[0], bounds[1], Nt, shuffle=False) (random.py:48-66)
            dba = (bounds[1] - bounds[0])/Nt
            r = bounds[0] + (np.arange(Nt) + np.random.default_rng().uniform(0, 1, Nt))*dba
            if dist.world() > 1:         # every rank must evaluate the same positions (collectives inside the cost)
                r = dist.broadcast_floats(r, engine.device())
            vals = np.array([self._focus_cost(L, float(z), method) for z in r])

        if method == "RMS Spot Size":
            # direct solution (raytracer.py:1422-1447), extended by ray weights
            m0, m1 = L.moments(0, [bounds[0]]), L.moments(0, [bounds[1]])
            pb0, pb1 = m0[2:4]/m0[0], m1[2:4]/m1[0]
            vz = bounds[1] - bounds[0]
            vx, vy = pb1[0] - pb0[0], pb1[1] - pb0[1]
            dnorm, num, _, _ = L.moments(2, [0.0, pb0[0], pb0[1], vx/vz, vy/vz])
            d = -num/dnorm if dnorm else np.mean(bounds)
            res = scipy.optimize.OptimizeResult()
            res.x = float(np.clip(d, bounds[0], bounds[1]))
            res.fun = self._focus_cost(L, res.x, "RMS Spot Size")
        else:
            cost = lambda z, method: self._focus_cost(L, float(z[0]), method)  # noqa: E731
            if method == "Irradiance Variance":
                res = scipy.optimize.minimize(cost, np.mean(bounds), args=method, tol=None, callback=None,
                                              options={'maxiter': 100}, bounds=[bounds], method="Nelder-Mead")
            else:
                res = scipy.optimize.minimize(cost, r[int(np.argmin(vals))], args=method, tol=None, callback=None,
                                              options={'maxiter': 30}, bounds=[bounds], method="COBYLA")
            res.x = float(res.x[0])

        rrl = (res.x - bounds[0]) < 10*(bounds[1] - bounds[0])/Nt
        rrr = (bounds[1] - res.x) < 10*(bounds[1] - bounds[0])/Nt
        if rrl or rrr:
            warning("Found minimum near search bounds, this can mean the focus is outside of the search range.")
        m = L.moments(0, [res.x])
        pos = (float(m[2]/m[0]), float(m[3]/m[0]), float(res.x))      # np.average(pa + sb*res.x, weights) with z = res.x
        if not return_cost:
            r = vals = None
        return res, dict(pos=pos, bounds=bounds, z=r, cost=vals, N=N_use)

    # -- iterative render (raytracer.py:1134-1279) ---------------------------------------------------------
    def iterative_render(self, N, detector_index=0, limit=None, projection_method="Equidistant", pos=None, extent=None):
        if not self.ray_sources:
            raise RuntimeError("Ray Source(s) Missing.")
        self.finish_trace()
        if not self.detectors:
            raise RuntimeError("Detector(s) Missing.")
        if (N := int(N)) <= 0:
            raise ValueError(f"Ray number N_rays needs to be a positive int, but is {N}.")
        if pos is None:
            if isinstance(detector_index, list):
                raise ValueError("detector_index list needs to have the same length as pos list")
            pos = [self.detectors[detector_index].pos]
        elif isinstance(pos, list) and not isinstance(pos[0], (list, np.ndarray)):
            pos = [pos]

        def expand(v, name, scalar_list=False):
            if not isinstance(v, list) or (scalar_list and isinstance(v[0], (int, float))):
                return [v]*len(pos)
            if len(v) != len(pos):
                raise ValueError(f"{name} list needs to have the same length as pos list")
            return v

        detector_index = expand(detector_index, "detector_index")
        limit = expand(limit, "limit")
        projection_method = expand(projection_method, "projection_method")
        extentc = list(expand(extent, "extent", scalar_list=True))

        rays_step = self.ITER_RAYS_STEP
        iterations = max(1, int(N/rays_step))
        if self._pretrace_check(rays_step):
            raise RuntimeError("Geometry checks failed. Tracing aborted. Check the warnings.")
        engine.ensure_init()
        torch = engine._torch()
        scene = self._scene_handle()
        nt = scene.nt
        msgs_cum = np.zeros((len(self.INFOS), nt), dtype=int)
        msgs_dev = None
        status_dev = torch.zeros(1, dtype=torch.int32, device=engine.device())      # OR-ed by every chunk, checked once
        powers = [rs.power for rs in self.ray_sources]
        nd = len(pos)
        images = [None]*nd
        grids = [None]*nd
        scratch = [None]*nd

        def det_records():
            recs = []
            for j in range(nd):
                if extentc[j] is not None and not isinstance(extentc[j], (list, np.ndarray)):
                    raise ValueError(f"Invalid extent '{extentc[j]}'.")
                self.detectors[detector_index[j]].move_to(pos[j])
                recs.append(detector_record(self.detectors[detector_index[j]].surface, projection_method[j], extentc[j]))
            return recs

        # fused mode: every chunk is generated, traced, tested against all detector positions and binned in one
        # kernel launch per group of <= 8 detectors; nothing is stored per surface
        for i in range(iterations):
            if i == iterations - 1:
                rays_step += int(N - iterations*rays_step)
            N_list = dist.shared_split(rays_step, powers, engine.device())
            self._trace_count += 1
            seed = (int(self.seed) << 20) + self._trace_count
            rays = self._generated(scene, N_list, dist.shard_sources(N_list), 0, seed)
            if getattr(rays, "gen_status", None) is not None:
                status_dev.bitwise_or_(rays.gen_status)
            recs = det_records()
            if i == 0:
                # auto extents from the first chunk (raytracer.py:1042-1046, 1262): range pass without binning
                auto = [j for j in range(nd) if extentc[j] is None]
                for g0 in range(0, len(auto), 8):
                    grp = auto[g0:g0 + 8]
                    rng = engine.trace_render(scene, rays, [recs[j] for j in grp], status=status_dev)
                    rows = dist.gather_rows(rng.reshape(-1)).reshape(-1, len(grp), 4)    # (world, detectors, 4)
                    for k, j in enumerate(grp):
                        r = np.array([rows[:, k, 0].min(), rows[:, k, 1].max(), rows[:, k, 2].min(), rows[:, k, 3].max()])
                        e = self.detectors[detector_index[j]].pos[:2].repeat(2)
                        extentc[j] = r.copy() if r[0] <= r[1] else e
                recs = det_records()
                for j in range(nd):
                    proj = projection_method[j] if recs[j]["projection"] else None
                    det = self.detectors[detector_index[j]]
                    pname = f": {det.desc}" if det.desc != "" else ""
                    im = RenderImage(long_desc=f"{Detector.abbr}{detector_index[j]}{pname} at z = {det.pos[2]:.5g} mm",
                                     extent=np.array(extentc[j], dtype=np.float64), projection=proj)
                    im._limit = limit[j]
                    im._fix_extent()
                    grids[j] = im._grid()
                    Nx, Ny = grids[j]
                    im._data_dev = torch.zeros((Ny, Nx, 4), dtype=torch.float64, device=engine.device())
                    im._counts_dev = torch.zeros((Ny, Nx), dtype=torch.int32, device=engine.device())
                    scratch[j] = torch.zeros((Ny, Nx, 4), dtype=torch.float64, device=engine.device())
                    images[j] = im
            msgs = None
            for g0 in range(0, nd, 8):
                grp = list(range(g0, min(nd, g0 + 8)))
                for j in grp:
                    scratch[j].zero_()
                m = engine.trace_render(scene, rays, [recs[j] for j in grp], extents=[images[j].extent for j in grp],
                                        grids=[grids[j] for j in grp], imgs=[scratch[j] for j in grp],
                                        cnts=[images[j]._counts_dev for j in grp], status=status_dev)
                msgs = m if msgs is None else msgs
                for j in grp:
                    # Imi._data *= rays_step / N ; DIm_res[j]._data += Imi._data   (raytracer.py:1257-1264)
                    images[j]._data_dev.add_(scratch[j], alpha=rays_step/N)
            # messages are accumulated on the device: no host synchronisation per chunk, the next chunk's generation
            # and trace are queued while this one runs
            msgs_dev = msgs.clone() if msgs_dev is None else msgs_dev.add_(msgs)
        m, st = dist.reduce_msgs_status(msgs_dev, status_dev)
        msgs_cum += m
        engine.raise_status(st)
        if st & 16:
            # the fused kernel decides "ray starts behind the detector" from the first point, which presumes z never
            # decreases along a ray (true for every ray the tracer produces: s_z > 0 is enforced); reported, not hidden
            raise RuntimeError("A ray position moved backwards in z during an iterative render; "
                               "use trace() + detector_image() for this geometry.")
        for j in range(nd):
            dist.allreduce_sum_(images[j]._data_dev)
            dist.allreduce_sum_(images[j]._counts_dev)
        # the filter is applied to the finished images (raytracer.py:1268-1272)
        for j in range(nd):
            if limit[j] is not None:
                images[j]._apply_rayleigh_filter()
        self._msgs = msgs_cum
        self._show_messages(N)
        return images
