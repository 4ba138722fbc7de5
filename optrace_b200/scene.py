"""Scene flattening: the Python object graph of a Raytracer -> the POD records of include/otb.h.

`FlatScene` is the single description both consumers read:
  * the CUDA engine, through `FlatScene.to_ctypes()` -> OtbSceneDesc (C-ABI), and
  * the test oracle (oracle/), which reads the same dict records — so the parity tests also cover
    the flattening itself.

The step list reproduces the element walk of Raytracer.trace (raytracer.py:274-278, 307-397, 492-508):
z-sorted Lens / Filter / Aperture elements plus the invisible end aperture at the outline's z-end,
a Lens contributing two steps (front, back), an IdealLens one.
"""
from __future__ import annotations

import ctypes as C
import hashlib

import numpy as np

from . import _cabi
from .surfaces import Surface, RectangularSurface, RingSurface, SlitSurface, NPAR
from .elements import Lens, IdealLens, Filter, Aperture

ROLE_LENS_FRONT, ROLE_LENS_BACK, ROLE_IDEAL, ROLE_FILTER, ROLE_APERTURE = range(5)


class FlatScene:
    def __init__(self):
        self.surfaces: list[dict] = []
        self.steps: list[dict] = []
        self.media: list[dict] = []
        self.filters: list[dict] = []
        self.aux = np.zeros(0, dtype=np.float64)
        self.outline = [0.0]*6
        self.no_pol = False
        self.arithmetic = 0          # OTB_ARITH_EXACT / OTB_ARITH_RELAXED
        self.medium0 = 0
        self.n_hurb = 0
        self.hurb_factor = 2**0.5
        self.user_funcs: list = []      # (kind, callable, args) for the device-function transpiler
        self.source_func_ids: dict = {}  # id(RaySource) -> slot of its orientation function
        self._keep = None

    @property
    def nt(self) -> int:
        """number of stored sections per ray: tracing surfaces + end absorber + source point (raytracer.py:278)"""
        return len(self.steps) + 1

    # -- construction helpers ------------------------------------------------------------------
    def _add_aux(self, arr) -> tuple[int, int]:
        if arr is None:
            return 0, 0
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        off = self.aux.shape[0]
        self.aux = np.concatenate((self.aux, arr))
        return off, arr.shape[0]

    def _add_func(self, kind: str, fn, args) -> int:
        for i, (k, f, a) in enumerate(self.user_funcs):
            if k == kind and f is fn and a == args:
                return i
        self.user_funcs.append((kind, fn, args))
        return len(self.user_funcs) - 1

    def add_surface(self, surf: Surface) -> int:
        rec = surf._record()
        rec["aux_off"], _ = self._add_aux(rec.pop("aux"))
        funcs = rec.pop("funcs")
        rec["func_ids"] = [-1, -1, -1]
        if funcs is not None:
            oned = funcs["one_d"]
            rec["func_ids"][0] = self._add_func("surf1d" if oned else "surf2d", *funcs["func"])
            if funcs["mask"][0] is not None:
                rec["func_ids"][1] = self._add_func("mask1d" if oned else "mask2d", *funcs["mask"])
            if funcs["deriv"][0] is not None:
                rec["func_ids"][2] = self._add_func("deriv1d" if oned else "deriv2d", *funcs["deriv"])
            rec["func_id"] = rec["func_ids"][0]
            rec["par"][11], rec["par"][12] = float(rec["func_ids"][1]), float(rec["func_ids"][2])
        self.surfaces.append(rec)
        return len(self.surfaces) - 1

    def add_medium(self, ri) -> int:
        for i, (obj, _) in enumerate(self._media_objs):
            if obj is ri or obj == ri:
                return i
        rec = ri._record()
        rec["aux_off"], n = self._add_aux(rec.pop("aux"))
        rec["aux_n"] = n//2
        fn = rec.pop("func")
        rec["func_id"] = self._add_func("wl", *fn) if fn is not None else -1
        self.media.append(rec)
        self._media_objs.append((ri, rec))
        return len(self.media) - 1

    def add_filter(self, spec) -> int:
        rec = spec._record()
        rec["aux_off"], n = self._add_aux(rec.pop("aux"))
        rec["aux_n"] = n//2
        fn = rec.pop("func")
        rec["func_id"] = self._add_func("wl", *fn) if fn is not None else -1
        self.filters.append(rec)
        return len(self.filters) - 1

    # -- ctypes view -----------------------------------------------------------------------------
    def to_ctypes(self) -> _cabi.OtbSceneDesc:
        ns, nst, nm, nf = len(self.surfaces), len(self.steps), len(self.media), len(self.filters)
        S = (_cabi.OtbSurface*max(ns, 1))()
        for i, r in enumerate(self.surfaces):
            fill_surface(S[i], r)
        ST = (_cabi.OtbStep*max(nst, 1))()
        for i, r in enumerate(self.steps):
            ST[i].role, ST[i].surface, ST[i].medium_after = r["role"], r["surface"], r["medium_after"]
            ST[i].filter, ST[i].hurb, ST[i].hurb_slot, ST[i].D = r["filter"], r["hurb"], r["hurb_slot"], r["D"]
        M = (_cabi.OtbMedium*max(nm, 1))()
        for i, r in enumerate(self.media):
            fill_medium(M[i], r)
        F = (_cabi.OtbFilter*max(nf, 1))()
        for i, r in enumerate(self.filters):
            F[i].type, F[i].inverse, F[i].func_id = r["type"], r["inverse"], r["func_id"]
            F[i].aux_off, F[i].aux_n = r["aux_off"], r["aux_n"]
            F[i].c[:] = r["c"]
        aux = np.ascontiguousarray(self.aux if self.aux.shape[0] else np.zeros(1), dtype=np.float64)
        d = _cabi.OtbSceneDesc()
        d.abi_version = 2
        d.n_surfaces, d.n_steps, d.n_media, d.n_filters = ns, nst, nm, nf
        d.no_pol, d.medium0, d.n_hurb, d.n_aux = int(self.no_pol), self.medium0, self.n_hurb, self.aux.shape[0]
        d.outline[:] = self.outline
        d.hurb_factor = self.hurb_factor
        d.arithmetic = int(self.arithmetic)
        d.surfaces, d.steps = C.cast(S, C.POINTER(_cabi.OtbSurface)), C.cast(ST, C.POINTER(_cabi.OtbStep))
        d.media, d.filters = C.cast(M, C.POINTER(_cabi.OtbMedium)), C.cast(F, C.POINTER(_cabi.OtbFilter))
        d.aux = aux.ctypes.data_as(C.POINTER(C.c_double))
        self._keep = (S, ST, M, F, aux)
        return d

    def fingerprint(self) -> str:
        """structural hash of everything that influences a trace (replaces the crepr() snapshot of
        raytracer.py:141-179)."""
        h = hashlib.sha256()
        # repr() of floats round-trips exactly, so hashing the record dicts is as strict as hashing the structs
        h.update(repr((self.surfaces, self.steps, self.media, self.filters, self.outline, self.hurb_factor,
                       self.no_pol, self.medium0, self.arithmetic)).encode())
        h.update(self.aux.tobytes())
        h.update(repr([(k, id(f), sorted(a.items())) for k, f, a in self.user_funcs]).encode())
        return h.hexdigest()


def fill_surface(dst: _cabi.OtbSurface, r: dict) -> None:
    dst.kind, dst.flags, dst.func_id = r["kind"], r["flags"], r["func_id"]
    dst.aux_off, dst.aux_n0, dst.aux_n1 = r.get("aux_off", 0), r["aux_n0"], r["aux_n1"]
    dst.pos[:] = r["pos"]
    dst.r, dst.z_min, dst.z_max = r["r"], r["z_min"], r["z_max"]
    dst.par[:] = r["par"]


def fill_medium(dst: _cabi.OtbMedium, r: dict) -> None:
    dst.model, dst.func_id, dst.aux_off, dst.aux_n = r["model"], r["func_id"], r["aux_off"], r["aux_n"]
    dst.c[:] = r["c"]


def standalone_surface(surf: Surface):
    """(record, aux, user_funcs) of a single surface, for the array-evaluation entry points"""
    fs = FlatScene()
    fs._media_objs = []
    fs.add_surface(surf)
    return fs.surfaces[0], fs.aux, fs.user_funcs


def flatten_raytracer(rt) -> FlatScene:
    """Builds the step list exactly like Raytracer.trace walks its elements (raytracer.py:297-397)."""
    fs = FlatScene()
    fs._media_objs = []
    o = rt.outline
    fs.outline = [float(v) for v in o]
    fs.no_pol = bool(rt.no_pol)
    fs.arithmetic = 1 if getattr(rt, "arithmetic", "exact") == "relaxed" else 0
    fs.hurb_factor = float(rt.HURB_FACTOR)
    fs.medium0 = fs.add_medium(rt.n0)

    # end absorber (raytracer.py:500-502)
    end = Aperture(RectangularSurface(dim=[o[1] - o[0], o[3] - o[2]]), pos=[(o[1] + o[0])/2, (o[2] + o[3])/2, o[5]])
    elements = [el for el in rt.elements if isinstance(el, (Lens, Filter, Aperture))] + [end]

    def step(role, surf, medium_after=-1, filt=-1, hurb=0, D=0.0):
        slot = -1
        if hurb:
            slot = fs.n_hurb
            fs.n_hurb += 1
        fs.steps.append(dict(role=role, surface=fs.add_surface(surf), medium_after=medium_after,
                             filter=filt, hurb=hurb, hurb_slot=slot, D=float(D)))

    for en, el in enumerate(elements):
        if isinstance(el, Lens):
            n2 = el.n2 or rt.n0
            if not el.is_ideal:
                step(ROLE_LENS_FRONT, el.front, medium_after=fs.add_medium(el.n))
                step(ROLE_LENS_BACK, el.back, medium_after=fs.add_medium(n2))
            else:
                step(ROLE_IDEAL, el.front, medium_after=fs.add_medium(n2), D=el.D)
        elif isinstance(el, Filter):
            step(ROLE_FILTER, el.surface, filt=fs.add_filter(el.spectrum))
        else:
            bend = bool(rt.use_hurb) and en != len(elements) - 1
            if bend and not isinstance(el.surface, (RingSurface, SlitSurface)):
                raise ValueError(f"Ray bending for surface type {type(el.surface).__name__} not implemented.")
            step(ROLE_APERTURE, el.surface, hurb=int(bend))
    # orientation functions of the sources (ray_source.py:274-276) are compiled into the same engine variant as the
    # scene's surface / medium callables: the generator runs inside the trace kernels
    fs.source_func_ids = {}
    for rs in getattr(rt, "ray_sources", []):
        if getattr(rs, "orientation", None) == "Function" and callable(rs.or_func):
            fs.source_func_ids[id(rs)] = fs._add_func("orient", rs.or_func, dict(rs.or_args))
    return fs


def detector_record(det_surface: Surface, projection: str | None, extent) -> dict:
    """dict form of OtbDetector"""
    from .surfaces import SphericalSurface
    rec, aux, funcs = standalone_surface(det_surface)
    proj = 0
    if isinstance(det_surface, SphericalSurface) and projection is not None:
        names = ["Equidistant", "Orthographic", "Equal-Area", "Stereographic"]
        if projection not in names:
            raise ValueError(f"Invalid projection_method {projection}, must be one of {names}.")
        proj = 1 + names.index(projection)
    return dict(surface=rec, projection=proj, has_extent=int(extent is not None),
                extent=[float(v) for v in extent] if extent is not None else [0.0]*4,
                R=getattr(det_surface, "R", 0.0))


def fill_detector(dst: _cabi.OtbDetector, r: dict) -> None:
    fill_surface(dst.surface, r["surface"])
    dst.projection, dst.has_extent = r["projection"], r["has_extent"]
    dst.extent[:] = r["extent"]
    # sphere projections need R (spherical_surface.py:50-92): passed in the spare conic slot
    dst.surface.par[10] = float(r["R"])
