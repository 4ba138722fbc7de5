"""Scene elements of the drop-in API: Element, Lens, IdealLens, Filter, Aperture, Detector, RaySource, Group.

Host-side containers only (positions, thickness bookkeeping, z-ordering); constructor signatures and
semantics follow optrace/tracer/geometry/{element,lens,ideal_lens,filter,aperture,detector,ray_source,group}.py.
No per-ray work happens here: RaySource flattens itself into an `OtbSource` record plus inverse-CDF tables
for the Philox-based device generator (replacement of RaySource.create_rays, ray_source.py:204-437).
"""
from __future__ import annotations

import copy as _copy
from typing import Callable

import numpy as np

from .options import warning
from .surfaces import (Surface, Point, Line, CircularSurface, RectangularSurface, RingSurface, SlitSurface,
                       DataSurface2D, FunctionSurface2D, AsphericSurface, _Shape)
from .media import RefractionIndex, TransmissionSpectrum, LightSpectrum
from .images import RGBImage, GrayscaleImage
from . import color


class Element(_Shape):
    """geometry/element.py"""
    abbr = "EL"
    _allow_non_2D = True

    def __init__(self, front, pos, back=None, d1: float = None, d2: float = None, **kwargs):
        super().__init__(**kwargs)
        ok = (Surface, Point, Line) if self._allow_non_2D else (Surface,)
        if not isinstance(front, ok):
            raise TypeError(f"front needs to be of type {ok}, but is {type(front)}.")
        if back is not None and not isinstance(back, ok):
            raise TypeError(f"back needs to be of type {ok}, but is {type(back)}.")
        self._front = front.copy()
        self._back = back.copy() if back is not None else None
        for name, v in (("d1", d1), ("d2", d2)):
            if v is not None and not isinstance(v, (int, float)):
                raise TypeError(f"{name} needs to be a number.")
        self._d1 = float(d1) if d1 is not None else None
        self._d2 = float(d2) if d2 is not None else None
        if self.has_back():
            if d1 is None or d2 is None:
                raise ValueError("d1 and d2 need to be specified for a Element with a back surface")
            if d1 < 0 or d2 < 0:
                raise ValueError(f"Thicknesses d1, d2 need to be non-negative but are {d1=} and {d2=}.")
        self.move_to(pos)

    front = property(lambda self: self._front)
    back = property(lambda self: self._back)
    d1 = property(lambda self: self._d1)
    d2 = property(lambda self: self._d2)
    surface = property(lambda self: self._front)

    def has_back(self) -> bool:
        return self._back is not None

    def set_surface(self, surf: Surface) -> None:
        if self.has_back():
            raise RuntimeError("Replacing of Surfaces only supported for objects with one surface")
        pos = self._front.pos
        self._front = surf.copy()
        self._front.move_to(pos)

    def move_to(self, pos) -> None:
        if not isinstance(pos, (list, np.ndarray)):
            raise TypeError("pos needs to be a list or array.")
        pos = np.asarray_chkfinite(pos, dtype=np.float64)
        if pos.shape[0] != 3:
            raise ValueError("pos needs to have 3 elements.")
        if not self.has_back():
            self._front.move_to(pos)
        else:
            self._front.move_to(pos - [0, 0, self._d1])
            self._back.move_to(pos + [0, 0, self._d2])

    @property
    def pos(self) -> np.ndarray:
        return self._front.pos + [0, 0, 0 if not self.has_back() else self._d1]

    @property
    def extent(self):
        if not self.has_back():
            return self._front.extent
        exts = np.column_stack((self._front.extent, self._back.extent))
        ext = np.zeros(6)
        ext[[0, 2, 4]] = np.min(exts, axis=1)[[0, 2, 4]]
        ext[[1, 3, 5]] = np.max(exts, axis=1)[[1, 3, 5]]
        return tuple(ext)

    def get_desc(self, fallback: str = None) -> str:
        s1 = type(self._front).__name__
        fb = f"{s1} + {type(self._back).__name__}, z = {self.pos[2]:.04g}" if self.has_back() \
            else f"{s1}, z = {self.pos[2]:.04g}"
        return super().get_desc(fb)

    def flip(self) -> None:
        """element.py:179-198"""
        if self.has_back():
            self._back.flip()
            self._front.flip()
            zp = self.pos[2]
            self._front.move_to([*self._front.pos[:2], zp + self._d1])
            self._back.move_to([*self._back.pos[:2], zp - self._d2])
            self._front, self._back = self._back, self._front
            self._d1, self._d2 = self._d2, self._d1
        else:
            self._front.flip()

    def rotate(self, angle: float) -> None:
        self._front.rotate(angle)
        if self.has_back():
            self._back.rotate(angle)


class Lens(Element):
    """geometry/lens.py"""
    abbr = "L"
    _allow_non_2D = False
    is_ideal = False

    def __init__(self, front: Surface, back: Surface, n: RefractionIndex, pos, de: float = 0, d: float = None,
                 d1: float = None, d2: float = None, n2: RefractionIndex = None, **kwargs):
        if not isinstance(n, RefractionIndex):
            raise TypeError("n needs to be a RefractionIndex.")
        if n2 is not None and not isinstance(n2, RefractionIndex):
            raise TypeError("n2 needs to be a RefractionIndex or None.")
        self.n, self.n2 = n, n2
        d1 = float(d1) if d1 is not None else d1
        d2 = float(d2) if d2 is not None else d2
        if isinstance(front, Surface) and isinstance(back, Surface):
            # thickness modes, lens.py:53-85
            if d is not None:
                de = d - front.dp - back.dn
                if de < 0:
                    d1 = d/2
                    d2 = d/2
            if de is not None and d1 is None and d2 is None:
                if de < 0:
                    d1 = -de/2
                    d2 = -de/2
                else:
                    d1 = de/2. + front.dp
                    d2 = de/2. + back.dn
            elif d1 is None or d2 is None:
                raise ValueError("Both thicknesses d1, d2 need to be specified")
        super().__init__(front, pos, back, d1, d2, **kwargs)

    @property
    def d(self) -> float:
        return self.d1 + self.d2

    @property
    def de(self) -> float:
        return float(self.back.z_min - self.front.z_max)


class IdealLens(Lens):
    """geometry/ideal_lens.py"""
    is_ideal = True

    def __init__(self, r: float, D: float, pos, n2: RefractionIndex = None, **kwargs):
        if not isinstance(D, (int, float)):
            raise TypeError("D needs to be a number.")
        np.asarray_chkfinite(D)
        self.D = float(D)
        if not D:
            raise ValueError("Optical Power needs to be non-zero")
        super().__init__(front=CircularSurface(r=r), back=CircularSurface(r=r),
                         n=RefractionIndex("Constant", n=1), pos=pos, d=0, n2=n2, **kwargs)


class Filter(Element):
    """geometry/filter.py"""
    abbr = "F"
    _allow_non_2D = False

    def __init__(self, surface: Surface, pos, spectrum: TransmissionSpectrum, **kwargs):
        super().__init__(surface, pos, **kwargs)
        if not isinstance(spectrum, TransmissionSpectrum):
            raise TypeError("spectrum needs to be a TransmissionSpectrum.")
        self.spectrum = spectrum

    def __call__(self, wl):
        return self.spectrum(wl)


class Aperture(Element):
    """geometry/aperture.py"""
    abbr = "AP"
    _allow_non_2D = False

    def __init__(self, surface: Surface, pos, **kwargs):
        super().__init__(surface, pos, **kwargs)


class PointMarker(Element):
    """geometry/marker/point_marker.py: text / point annotation of a geometry (an Element on a Point); carried
    through Group bookkeeping (the .zmx importer labels its groups with one), ignored by the tracer"""
    abbr = "M"
    _allow_non_2D = True

    def __init__(self, desc: str, pos, text_factor: float = 1., marker_factor: float = 1., label_only: bool = False,
                 **kwargs):
        for nm, v in (("text_factor", text_factor), ("marker_factor", marker_factor)):
            if not isinstance(v, (int, float)) or isinstance(v, bool):
                raise TypeError(f"{nm} needs to be a number.")
        if not isinstance(label_only, bool):
            raise TypeError("label_only needs to be bool.")
        self.marker_factor, self.text_factor, self.label_only = marker_factor, text_factor, label_only
        super().__init__(Point(), pos, desc=desc, **kwargs)


class Detector(Element):
    """geometry/detector.py"""
    abbr = "DET"
    _allow_non_2D = False

    def __init__(self, surface: Surface, pos, **kwargs):
        if isinstance(surface, (DataSurface2D, FunctionSurface2D, AsphericSurface)):
            raise RuntimeError("Classes and subclasses of DataSurface1D, DataSurface2D, FunctionSurface2D"
                               " are not supported as Detector surfaces.")
        super().__init__(surface, pos, **kwargs)


class RaySource(Element):
    """geometry/ray_source.py"""
    divergences = ["None", "Lambertian", "Isotropic", "Function"]
    orientations = ["Constant", "Converging", "Function"]
    polarizations = ["Constant", "Uniform", "List", "Function", "x", "y", "xy"]
    abbr = "RS"
    _allow_non_2D = True
    _max_image_px = 2e6

    def __init__(self, surface, pos=None, divergence: str = "None", div_angle: float = 0.5, div_2d: bool = False,
                 div_axis_angle: float = 0, div_func: Callable = None, div_args: dict = {},
                 spectrum: LightSpectrum = None, power: float = 1., s=None, s_sph=None,
                 orientation: str = "Constant", conv_pos=None, or_func: Callable = None, or_args: dict = {},
                 polarization: str = "Uniform", pol_angle: float = 0., pol_angles=None, pol_probs=None,
                 pol_func: Callable = None, pol_args: dict = {}, **kwargs):
        if isinstance(surface, (RGBImage, GrayscaleImage)):
            if surface.shape[0]*surface.shape[1] > self._max_image_px:
                raise RuntimeError(f"For performance reasons only images with less than {self._max_image_px/1e6}"
                                   " megapixels are allowed.")
            surface_ = RectangularSurface(dim=surface.s)
            self._image = surface
            if isinstance(surface, RGBImage):
                lin = color.srgb_to_srgb_linear(surface._data)
                If = color.power_from_srgb_linear(lin).flatten()
            else:
                If = color.srgb_to_srgb_linear(surface._data).ravel()
            Ifs = If.sum()
            if Ifs <= 0:
                raise ValueError("Image can not be completely black.")
            self._pIf = 1/Ifs*If
        else:
            surface_ = surface
            self._image = None
            self._pIf = None
        if not isinstance(surface_, (CircularSurface, RectangularSurface, RingSurface, Point, Line)) \
                or isinstance(surface_, SlitSurface):
            raise ValueError("Invalid surface type for a RaySource: " + type(surface_).__name__)
        pos = pos if pos is not None else [0, 0, 0]
        super().__init__(surface_, pos, **kwargs)

        if not isinstance(power, (int, float)) or power <= 0:
            raise ValueError("power needs to be a positive number.")
        self.power = float(power)
        if spectrum is None:
            from .presets import light_spectrum as _ls
            spectrum = _ls.d65
        if not isinstance(spectrum, LightSpectrum):
            raise TypeError("spectrum needs to be a LightSpectrum.")
        self.spectrum = spectrum

        if polarization not in self.polarizations:
            raise ValueError(f"Invalid polarization '{polarization}', must be one of {self.polarizations}.")
        if divergence not in self.divergences:
            raise ValueError(f"Invalid divergence '{divergence}', must be one of {self.divergences}.")
        if orientation not in self.orientations:
            raise ValueError(f"Invalid orientation '{orientation}', must be one of {self.orientations}.")
        for name, f in (("div_func", div_func), ("or_func", or_func), ("pol_func", pol_func)):
            if f is not None and not callable(f):
                raise TypeError(f"{name} needs to be callable.")
        if not isinstance(div_2d, bool):
            raise TypeError("div_2d needs to be bool.")
        if not isinstance(div_angle, (int, float)) or div_angle <= 0:
            raise ValueError("div_angle needs to be a number above 0.")
        self.polarization, self.pol_angle, self.pol_func = polarization, float(pol_angle), pol_func
        self.pol_angles = None if pol_angles is None else list(pol_angles)
        self.pol_probs = None if pol_probs is None else list(pol_probs)
        self.pol_args = _copy.deepcopy(pol_args)
        self.divergence, self.div_angle, self.orientation = divergence, float(div_angle), orientation
        self.conv_pos = np.asarray_chkfinite(conv_pos if conv_pos is not None else [0, 0, 0], dtype=np.float64)
        self.or_func, self.or_args = or_func, _copy.deepcopy(or_args)
        if s_sph is None:
            sv = s if s is not None else [0, 0, 1]
        else:
            theta, phi = np.radians(s_sph[0]), np.radians(s_sph[1])
            sv = [np.sin(theta)*np.cos(phi), np.sin(theta)*np.sin(phi), np.cos(theta)]
        sv = np.asarray_chkfinite(sv, dtype=np.float64)
        if sv.shape[0] != 3:
            raise TypeError("s needs to have 3 elements.")
        if not sv[2] > 0:
            raise ValueError("s[2] needs to be above 0.")
        self.s = sv/np.linalg.norm(sv)     # ray_source.py __setattr__ normalises s
        self.div_axis_angle, self.div_func = float(div_axis_angle), div_func
        self.div_2d, self.div_args = div_2d, _copy.deepcopy(div_args)

    # -- flattening for the device generator ------------------------------------------------------
    def _generator_record(self) -> dict:
        """dict form of OtbSource (+ tables); see scene.flatten_sources."""
        import scipy.integrate
        sf = self.front
        rec = dict(tables={}, geom=[0.0]*8, extent=[0.0]*4, img_w=0, img_h=0)
        if self._image is not None:
            rec["shape"] = 5 if isinstance(self._image, RGBImage) else 6
            rec["img_h"], rec["img_w"] = self._image.shape[:2]
            rec["extent"] = [float(v) for v in sf.extent[:4]]
            f = self._pIf
            keep = f > 0
            # discrete inverse CDF over pixels (random.py:129-140): zero-probability pixels are excluded
            rec["tables"]["pix_idx"] = np.nonzero(keep)[0].astype(np.float64)
            rec["tables"]["pix_cdf"] = np.cumsum(f[keep])
            if isinstance(self._image, RGBImage):
                lin = color.srgb_to_srgb_linear(self._image._data.reshape(-1, 3))
                lin[:, 0] *= color.SRGB_R_PRIMARY_POWER_FACTOR
                lin[:, 2] *= color.SRGB_B_PRIMARY_POWER_FACTOR
                cs = np.cumsum(lin, axis=-1)
                last = cs[:, -1, np.newaxis]
                cs /= np.where(last, last, 1)
                rec["tables"]["pix_rgb"] = np.ascontiguousarray(cs[:, :2]).ravel()  # thresholds r, r+g
        elif isinstance(sf, Point):
            rec["shape"] = 0
        elif isinstance(sf, Line):
            ang = np.deg2rad(sf.angle)
            rec["shape"], rec["geom"][:3] = 1, [sf.r, float(np.cos(ang)), float(np.sin(ang))]
        elif isinstance(sf, RingSurface):
            rec["shape"], rec["geom"][:2] = 3, [sf.ri, sf.r]
        elif isinstance(sf, CircularSurface):
            rec["shape"], rec["geom"][:2] = 2, [0.0, sf.r]
        elif isinstance(sf, RectangularSurface):
            rec["shape"] = 4
            rec["geom"][:4] = [float(sf.dim[0]), float(sf.dim[1]), float(np.cos(sf._angle)), float(np.sin(sf._angle))]
            rec["geom"][4] = 1.0 if sf._angle else 0.0
        rec["pos"] = [float(v) for v in sf.pos]

        if self.orientation == "Function" and not callable(self.or_func):
            raise TypeError("RaySource.or_func needs to be callable.")
        # "Function": or_func(x, y) is translated into a device function (userfunc.py, kind "orient"); the scene
        # flattening registers it and fills in the slot (scene.flatten_raytracer)
        rec["orientation"] = ["Constant", "Converging", "Function"].index(self.orientation)
        rec["s"] = [float(v) for v in self.s]
        rec["conv_pos"] = [float(v) for v in self.conv_pos]

        rec["divergence"] = self.divergences.index(self.divergence)
        rec["div_2d"] = int(self.div_2d)
        rec["div_sin"] = float(np.sin(np.radians(self.div_angle)))
        rec["div_angle"] = float(np.radians(self.div_angle))
        rec["div_axis"] = float(np.radians(self.div_axis_angle))
        if self.divergence == "Function":
            if not callable(self.div_func):
                raise TypeError("RaySource.div_func needs to be callable.")
            x = np.linspace(0, np.radians(self.div_angle), 1000)
            f = self.div_func(x, **self.div_args)*(np.sin(x) if not self.div_2d else 1.0)
            rec["tables"]["div"] = (x, scipy.integrate.cumulative_trapezoid(f, initial=0))

        pol = self.polarization
        rec["pol_angle"] = 0.0
        if pol in ("x", "y", "Constant"):
            rec["polarization"] = 0
            rec["pol_angle"] = {"x": 0., "y": float(np.pi/2)}.get(pol, float(np.radians(self.pol_angle)))
        elif pol == "Uniform":
            rec["polarization"] = 1
        elif pol in ("xy", "List"):
            rec["polarization"] = 2
            if pol == "xy":
                ang, pr = np.array([0, np.pi/2]), np.ones(2)
            else:
                if self.pol_angles is None:
                    raise TypeError("RaySource.pol_angles needs to be a list.")
                pr = np.ones_like(self.pol_angles, dtype=np.float64) if self.pol_probs is None \
                    else np.asarray(self.pol_probs, dtype=np.float64)
                ang = np.radians(np.asarray(self.pol_angles, dtype=np.float64))
            keep = pr > 0
            rec["tables"]["pol"] = (ang[keep], np.cumsum(pr[keep]))
        else:  # Function
            rec["polarization"] = 3
            if not callable(self.pol_func):
                raise TypeError("RaySource.pol_func needs to be callable.")
            x = np.linspace(0, 2*np.pi, 5000)
            f = self.pol_func(x, **self.pol_args)
            rec["tables"]["pol"] = (x, scipy.integrate.cumulative_trapezoid(f, initial=0))

        if rec["shape"] == 5:
            rec["wl"] = dict(mode=5, wl=[0, 0, 0, 0], tab=None)
        else:
            rec["wl"] = self.spectrum._sampling_record()
        rec["power"] = self.power
        return rec


class Group(_Shape):
    """geometry/group.py (markers and volumes are GUI decoration and not modelled here)"""

    def __init__(self, elements: list = None, n0: RefractionIndex = None, **kwargs):
        super().__init__(**kwargs)
        self.lenses, self.apertures, self.filters = [], [], []
        self.detectors, self.ray_sources = [], []
        self.markers, self.volumes = [], []
        self.n0 = n0
        if elements is not None:
            self.add(elements)

    @property
    def n0(self):
        return self._n0

    @n0.setter
    def n0(self, val):
        if val is None:
            val = RefractionIndex("Constant", n=1)
        if not isinstance(val, RefractionIndex):
            raise TypeError("n0 needs to be a RefractionIndex.")
        self._n0 = val

    @property
    def _elements(self) -> list:
        return [*self.lenses, *self.apertures, *self.filters, *self.ray_sources, *self.detectors,
                *self.markers, *self.volumes]

    @property
    def elements(self) -> list:
        return sorted(self._elements, key=lambda el: el.pos[2])

    @property
    def pos(self):
        return self.elements[0].pos if len(self._elements) else [0, 0, 0]

    @property
    def tracing_surfaces(self) -> list:
        """group.py:78-95"""
        surfs = []
        for el in self.elements:
            if isinstance(el, (Lens, Filter, Aperture)):
                surfs.append(el.front)
                if el.has_back() and not isinstance(el, IdealLens):
                    surfs.append(el.back)
        return surfs

    @property
    def extent(self):
        els = self._elements
        if not len(els):
            return 0, 0, 0, 0, 0, 0
        ext = np.array([np.array(el.extent) for el in els])
        mx, mn = np.max(ext, axis=0), np.min(ext, axis=0)
        return mn[0], mx[1], mn[2], mx[3], mn[4], mx[5]

    def move_to(self, pos) -> None:
        pos = np.asarray_chkfinite(pos, dtype=np.float64)
        if pos.shape[0] != 3:
            raise ValueError("pos needs to have exactly 3 elements.")
        pos0 = self.pos
        for el in self._elements:
            el.move_to(el.pos - (pos0 - pos))

    def flip(self, y0: float = 0, z0: float = None) -> None:
        """group.py:152-196"""
        if not len(self._elements):
            return
        els = self.elements
        ns = [self.n0] + [L.n2 for L in els if isinstance(L, Lens)]
        z0 = np.mean(self.extent[4:]) if z0 is None else z0
        self.clear()
        els.reverse()
        self.add(els)
        for el in els:
            el.flip()
            el.move_to([el.pos[0], y0 - (el.pos[1] - y0), z0 - (el.pos[2] - z0)])
        ns.reverse()
        ns = [n if n is not None else self.n0 for n in ns]
        self.n0 = ns[0]
        for n2, L in zip(ns[1:], self.lenses):
            L.n2 = n2

    def rotate(self, angle: float, x0: float = 0, y0: float = 0) -> None:
        if not len(self._elements):
            return
        ang = np.deg2rad(angle)
        for el in self.elements:
            xr, yr = el.pos[0] - x0, el.pos[1] - y0
            posr = [x0 + xr*np.cos(ang) - yr*np.sin(ang), y0 + xr*np.sin(ang) + yr*np.cos(ang), el.pos[2]]
            el.rotate(angle)
            el.move_to(posr)

    def add(self, el) -> None:
        if not isinstance(el, list) and self.has(el):
            warning("Element already included in geometry. Make a copy to include it another time.")
            return
        if isinstance(el, Aperture):
            self.apertures.append(el)
        elif isinstance(el, Filter):
            self.filters.append(el)
        elif isinstance(el, RaySource):
            self.ray_sources.append(el)
        elif isinstance(el, Detector):
            self.detectors.append(el)
        elif isinstance(el, Lens):
            self.lenses.append(el)
        elif isinstance(el, PointMarker):
            self.markers.append(el)
        elif isinstance(el, Group):
            if self.n0 != el.n0:
                warning("Overwriting ambient index with index from new Group.")
                self.n0 = el.n0
            for e in el.elements:
                self.add(e)
        elif isinstance(el, list):
            for e in el:
                self.add(e)
        else:
            raise TypeError(f"Unsupported element type {type(el).__name__}.")

    def remove(self, el) -> bool:
        ok = False
        if isinstance(el, list):
            for e in el.copy():
                ok = self.remove(e) or ok
        elif isinstance(el, Group):
            for e in el._elements.copy():
                ok = self.remove(e) or ok
        else:
            for lst in (self.lenses, self.apertures, self.detectors, self.volumes, self.filters,
                        self.ray_sources, self.markers):
                for l in lst.copy():
                    if l is el:
                        lst.remove(l)
                        ok = True
        return ok

    def has(self, el) -> bool:
        return any(e is el for e in self._elements)

    def clear(self) -> None:
        for lst in (self.lenses, self.apertures, self.filters, self.detectors, self.ray_sources,
                    self.markers, self.volumes):
            lst[:] = []
