"""Surface model of the drop-in API (host side).

The classes keep the constructor signatures and geometric bookkeeping (pos, r, z_min, z_max,
extent, move_to/flip/rotate/copy) of optrace/tracer/geometry/surface/*.py, but they do no
per-ray work on the host: every surface flattens itself into an `OtbSurface` record
(include/otb.h) that the CUDA engine consumes, and `find_hit` / `normals` on arrays are
evaluated by the CUDA library (no CPU fallback).

Host numpy code in this file is scene SETUP only (z-bounds of user functions, spline fits,
mask tests used by geometry checks) — the same work the reference does once per surface
construction, never per traced ray.

Scalars that enter the device formulas (rho, (k+1)*rho**2, cos(angle), finite-difference
step, ...) are computed HERE with the same Python/numpy scalar expressions the reference
uses inside its vectorised formulas, so the device sees bit-identical constants.
"""
from __future__ import annotations

import copy as _copy
from typing import Callable

import numpy as np

from .options import warning
from . import _state

C_EPS = 1e-6    # Surface.C_EPS (surface.py:17)
N_EPS = 1e-10   # Surface.N_EPS (surface.py:20)

# OtbSurfKind
K_CIRCLE, K_RECT, K_RING, K_SLIT, K_CONIC, K_TILTED, K_ASPHERE, K_FUNC, K_DATA = range(9)
# OTB_SF_*
F_ROTATED, F_1D, F_HAS_DERIV, F_HAS_MASK, F_FLAT, F_ROTSYM = 1, 2, 4, 8, 16, 32
NPAR = 20


def _rot(x, y, alpha):
    """Surface._rotate_rc (surface.py:427-434): identity when alpha == 0."""
    if alpha:
        return x*np.cos(alpha) - y*np.sin(alpha), x*np.sin(alpha) + y*np.cos(alpha)
    return x, y


class _Shape:
    """Common base of surfaces, points and lines (position + copy)."""

    def __init__(self, desc: str = "", long_desc: str = ""):
        self.desc = desc
        self.long_desc = long_desc

    def __setattr__(self, key, val):
        object.__setattr__(self, key, val)
        _state.EPOCH[0] += 1        # scene epoch: see _state.py

    def copy(self):
        return _copy.deepcopy(self)

    def get_desc(self, fallback: str = "") -> str:
        return self.desc if self.desc != "" else fallback

    def get_long_desc(self, fallback: str = "") -> str:
        return self.long_desc if self.long_desc != "" else self.get_desc(fallback)


class Point(_Shape):
    """geometry/point.py"""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.pos = np.array([0., 0., 0.])
        self.z_min = self.z_max = 0.

    def move_to(self, pos):
        self.pos = np.asarray_chkfinite(pos, dtype=np.float64)
        self.z_min = self.z_max = float(pos[2])

    def flip(self):
        pass

    def rotate(self, angle):
        pass

    @property
    def extent(self):
        return tuple(self.pos.repeat(2))


class Line(_Shape):
    """geometry/line.py"""

    def __init__(self, r: float, angle: float = 0, **kwargs):
        super().__init__(**kwargs)
        if not isinstance(r, (int, float)) or not isinstance(angle, (int, float)):
            raise TypeError("r and angle need to be numbers.")
        if r <= 0:
            raise ValueError("r needs to be above 0.")
        self.pos = np.array([0., 0., 0.])
        self.r = float(r)
        self.angle = float(angle)
        self.z_min = self.z_max = 0.

    def move_to(self, pos):
        self.pos = np.asarray_chkfinite(pos, dtype=np.float64)
        self.z_min = self.z_max = float(pos[2])

    def flip(self):
        self.angle *= -1

    def rotate(self, angle):
        self.angle += angle

    @property
    def extent(self):
        ang = np.deg2rad(self.angle)
        return (self.pos[0] - self.r*np.cos(ang), self.pos[0] + self.r*np.cos(ang),
                self.pos[1] - self.r*np.sin(ang), self.pos[1] + self.r*np.sin(ang), self.z_min, self.z_max)


class Surface(_Shape):
    """Base surface (surface.py:15-497)."""

    C_EPS = C_EPS
    N_EPS = N_EPS
    rotational_symmetry = False
    _kind = K_CIRCLE

    def __init__(self, r: float, **kwargs):
        super().__init__(**kwargs)
        if not isinstance(r, (int, float)):
            raise TypeError("r needs to be a number.")
        if r <= 0:
            raise ValueError("r needs to be above 0.")
        self.pos = np.array([0., 0., 0.])
        self.r = float(r)
        self.parax_roc = None
        self.z_min, self.z_max = np.nan, np.nan

    # -- bookkeeping --------------------------------------------------------------------------
    def is_flat(self) -> bool:
        return self.z_max == self.z_min

    def move_to(self, pos) -> None:
        """surface.py:95-110"""
        self.z_min += pos[2] - self.pos[2]
        self.z_max += pos[2] - self.pos[2]
        self.pos = np.asarray_chkfinite(pos, dtype=np.float64)

    @property
    def extent(self):
        """surface.py:112-120"""
        return (*(self.r*np.array([-1, 1, -1, 1]) + self.pos[:2].repeat(2)), self.z_min, self.z_max)

    @property
    def ds(self) -> float:
        return float(self.z_max - self.z_min)

    @property
    def dn(self) -> float:
        return float(self.pos[2] - self.z_min)

    @property
    def dp(self) -> float:
        return float(self.z_max - self.pos[2])

    def flip(self) -> None:
        assert self.is_flat()

    def rotate(self, angle: float) -> None:
        assert self.rotational_symmetry

    @property
    def info(self) -> str:
        return (f"{type(self).__name__}, pos = [{self.pos[0]:.5g} mm, {self.pos[1]:.5g} mm, "
                f"{self.pos[2]:.5g} mm], r = {self.r:.5g} mm")

    # -- host-side setup helpers (not on the ray path) ----------------------------------------
    def _values(self, x, y):
        """relative height without masking; flat by default"""
        return np.broadcast_to(0., np.shape(x))

    def mask(self, x, y):
        """surface.py:235-245 (absolute coordinates)"""
        x0, y0 = self.pos[0], self.pos[1]
        return (x - x0)**2 + (y - y0)**2 <= (self.r + self.N_EPS)**2

    def values(self, x, y):
        """surface.py:137-164: heights with the radially continued edge (setup/checks only)."""
        x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        if self.is_flat():
            return np.broadcast_to(self.z_max, x.shape)
        z = np.full_like(x, self.z_max, dtype=np.float64)
        inside = self.mask(x, y)
        z[inside] = self.pos[2] + self._values(x[inside] - self.pos[0], y[inside] - self.pos[1])
        r = self.r - self.N_EPS
        if np.any(~inside):
            if not self.rotational_symmetry:
                phi = np.arctan2(y[~inside] - self.pos[1], x[~inside] - self.pos[0])
                z[~inside] = self.pos[2] + self._values(r*np.cos(phi), r*np.sin(phi))
            else:
                z[~inside] = self.pos[2] + self._values(np.array([r]), np.array([0.]))[0]
        return z

    def edge(self, nc: int):
        """surface.py:287-304"""
        theta = np.linspace(-3/4*np.pi, 5/4*np.pi, nc)
        xd, yd = self.r*np.cos(theta), self.r*np.sin(theta)
        return xd + self.pos[0], yd + self.pos[1], self._values(xd, yd) + self.pos[2]

    def _find_bounds(self):
        """surface.py:57-93: sunflower + edge sampling of z_min/z_max (setup)."""
        N = 50000
        ind = np.arange(0, N, dtype=np.float64)
        r = np.sqrt(ind/N)*self.r
        phi = 2*np.pi*(1 + 5**0.5)/2*ind
        rcos, rsin = r*np.cos(phi), r*np.sin(phi)
        vals = np.array(self._values(rcos, rsin), dtype=np.float64)
        m = self.mask(rcos - self.pos[0], rsin - self.pos[1])
        vals[~m] = np.nan
        xv, yv, vals2 = self.edge(3001)
        vals2 = vals2 - self.pos[2]
        m = self.mask(xv, yv)
        vals2[~m] = np.nan
        return (float(min(np.nanmin(vals), np.nanmin(vals2))), float(max(np.nanmax(vals), np.nanmax(vals2))))

    def _fd_eps(self) -> float:
        """finite-difference step of Surface.normals (surface.py:266-270)"""
        eps_f = np.finfo(np.float64).eps
        eps_deriv = (3*eps_f*50)**(1/3)
        ext = np.array(self.extent)
        eps_num = np.spacing(ext[1::2] - ext[::2])
        return float(max(eps_deriv, *eps_num))

    def _edge_z(self) -> float:
        """relative height of the radially continued edge for rotationally symmetric surfaces
        (surface.py:153, 162)."""
        if self.is_flat():
            return 0.0
        r = self.r - self.N_EPS
        return float(self._values(np.array([r]), np.array([0.]))[0])

    # -- flattening ---------------------------------------------------------------------------
    def _base_record(self) -> dict:
        flags = (F_FLAT if self.is_flat() else 0) | (F_ROTSYM if self.rotational_symmetry else 0)
        return dict(kind=self._kind, flags=flags, func_id=-1, aux=None, aux_n0=0, aux_n1=0,
                    pos=[float(v) for v in self.pos], r=float(self.r),
                    z_min=float(self.z_min), z_max=float(self.z_max), par=[0.0]*NPAR,
                    funcs=None)

    def _record(self) -> dict:
        """Flattened description (dict form of OtbSurface, see scene.py)."""
        return self._base_record()

    # -- array evaluation on the device ---------------------------------------------------------
    def find_hit(self, p: np.ndarray, s: np.ndarray, where=None):
        """Surface.find_hit on arrays, evaluated by the CUDA engine (surface.py:307-414 and overrides).
        Returns (p_hit, is_hit, ill) like the reference."""
        from . import engine
        ph, hit, ill = engine.surface_find_hit(self, p, s)
        w_ = where if where is not None else slice(None)
        return ph[w_], hit[w_], ill[w_]

    def normals(self, x: np.ndarray, y: np.ndarray) -> np.ndarray:
        """Surface.normals on arrays, evaluated by the CUDA engine (surface.py:247-285 and overrides)."""
        from . import engine
        return engine.surface_normals(self, x, y)


class CircularSurface(Surface):
    """circular_surface.py"""
    rotational_symmetry = True
    _kind = K_CIRCLE

    def __init__(self, r: float, **kwargs):
        super().__init__(r, **kwargs)
        self.parax_roc = np.inf
        self.z_min = self.z_max = self.pos[2]


class RingSurface(Surface):
    """ring_surface.py"""
    rotational_symmetry = True
    _kind = K_RING

    def __init__(self, r: float, ri: float, **kwargs):
        super().__init__(r, **kwargs)
        if not isinstance(ri, (int, float)):
            raise TypeError("ri needs to be a number.")
        if ri <= 0:
            raise ValueError("ri needs to be above 0.")
        self.ri = float(ri)
        self.parax_roc = np.inf
        self.z_min = self.z_max = self.pos[2]
        if ri >= r:
            raise ValueError("ri needs to be smaller than r.")

    def mask(self, x, y):
        """ring_surface.py:123-133"""
        x0, y0 = self.pos[0], self.pos[1]
        r2 = (x - x0)**2 + (y - y0)**2
        return ((self.ri - self.N_EPS)**2 <= r2) & (r2 <= (self.r + self.N_EPS)**2)

    def _record(self):
        rec = self._base_record()
        rec["par"][0] = self.ri
        return rec


class RectangularSurface(Surface):
    """rectangular_surface.py"""
    rotational_symmetry = False
    _kind = K_RECT

    def __init__(self, dim, **kwargs):
        self._angle = 0
        super().__init__(1, **kwargs)
        if not isinstance(dim, (list, np.ndarray)):
            raise TypeError("dim needs to be a list or array.")
        dim = np.asarray_chkfinite(dim, dtype=np.float64)
        if dim.ndim != 1 or dim.shape[0] != 2:
            raise TypeError("dim needs to have two elements.")
        if dim[0] <= 0 or dim[1] <= 0:
            raise ValueError("Dimensions dim need to be positive.")
        self.dim = dim
        self.parax_roc = np.inf
        self.z_min = self.z_max = self.pos[2]

    @property
    def extent(self):
        """rectangular_surface.py:45-57"""
        sx = np.abs(self.dim[0]*np.cos(self._angle)) + np.abs(self.dim[1]*np.sin(self._angle))
        sy = np.abs(self.dim[0]*np.sin(self._angle)) + np.abs(self.dim[1]*np.cos(self._angle))
        return (self.pos[0] - sx/2, self.pos[0] + sx/2, self.pos[1] - sy/2, self.pos[1] + sy/2,
                self.z_min, self.z_max)

    @property
    def _extent(self):
        return -self.dim[0]/2, self.dim[0]/2, -self.dim[1]/2, self.dim[1]/2, 0., 0.

    def rotate(self, angle: float) -> None:
        self._angle += np.deg2rad(angle)

    def flip(self) -> None:
        self._angle *= -1

    def mask(self, x, y):
        """rectangular_surface.py:100-112"""
        xr, yr = _rot(x - self.pos[0], y - self.pos[1], -self._angle)
        xs, xe, ys, ye = self._extent[:4]
        return (xs - self.N_EPS <= xr) & (xr <= xe + self.N_EPS) & (ys - self.N_EPS <= yr) & (yr <= ye + self.N_EPS)

    def _rot_par(self, rec):
        a = self._angle
        if a:
            rec["flags"] |= F_ROTATED
        rec["par"][2], rec["par"][3] = float(np.cos(-a)), float(np.sin(-a))
        rec["par"][4], rec["par"][5] = float(np.cos(a)), float(np.sin(a))

    def _record(self):
        rec = self._base_record()
        rec["par"][0], rec["par"][1] = float(self.dim[0]), float(self.dim[1])
        self._rot_par(rec)
        return rec


class SlitSurface(RectangularSurface):
    """slit_surface.py"""
    _kind = K_SLIT

    def __init__(self, dim, dimi, **kwargs):
        super().__init__(dim, **kwargs)
        if not isinstance(dimi, (list, np.ndarray)):
            raise TypeError("dimi needs to be a list or array.")
        dimi = np.asarray_chkfinite(dimi, dtype=np.float64)
        if dimi.ndim != 1 or dimi.shape[0] != 2:
            raise TypeError("dimi needs to have two elements.")
        if dimi[0] >= self.dim[0] or dimi[1] >= self.dim[1]:
            raise ValueError("Dimensions dimi must be smaller than dimension dim.")
        if dimi[0] <= 0 or dimi[1] <= 0:
            raise ValueError("Dimensions dimi need to be positive.")
        self.dimi = dimi

    def mask(self, x, y):
        """slit_surface.py:89-102"""
        xr, yr = _rot(x - self.pos[0], y - self.pos[1], -self._angle)
        xs, xe, ys, ye = -self.dimi[0]/2, self.dimi[0]/2, -self.dimi[1]/2, self.dimi[1]/2
        inside = (xs + self.N_EPS <= xr) & (xr <= xe - self.N_EPS) & (ys + self.N_EPS <= yr) & (yr <= ye - self.N_EPS)
        return super().mask(x, y) & (~inside)

    def _record(self):
        rec = super()._record()
        rec["par"][6], rec["par"][7] = float(self.dimi[0]), float(self.dimi[1])
        return rec


class ConicSurface(Surface):
    """conic_surface.py"""
    rotational_symmetry = True
    _kind = K_CONIC

    def __init__(self, r: float, R: float, k: float, **kwargs):
        super().__init__(r, **kwargs)
        for name, v in (("R", R), ("k", k)):
            if not isinstance(v, (int, float)):
                raise TypeError(f"{name} needs to be a number.")
        R, k = float(R), float(k)
        if R == 0 or not np.isfinite(R):
            raise ValueError("R needs to be non-zero and finite. Use planar surface types for planar surfaces.")
        self.R, self.k = R, k
        self.parax_roc = R
        if (self.k + 1)*(self.r/self.R)**2 >= 1:
            raise ValueError("Surface radius r larger than radius of conic section.")
        z0 = self.pos[2]
        z1 = z0 + self._values(np.array([r]), np.array([0]))[0]
        self.z_min, self.z_max = min(z0, z1), max(z0, z1)

    @property
    def info(self):
        return super().info + f", R = {self.R:.5g} mm, k = {self.k:.5g}"

    def _values(self, x, y):
        """conic_surface.py:57-68"""
        k, rho = self.k, 1/self.R
        r2 = x**2 + y**2
        return rho*r2/(1 + np.sqrt(1 - (k+1)*rho**2*r2))

    def flip(self) -> None:
        """conic_surface.py:205-214"""
        self.R *= -1
        self.parax_roc *= -1
        a = self.pos[2] - (self.z_max - self.pos[2])
        b = self.pos[2] + (self.pos[2] - self.z_min)
        self.z_min, self.z_max = a, b

    def _conic_par(self, rec):
        k, rho = self.k, 1/self.R
        p = rec["par"]
        p[0], p[1], p[2] = k, rho, k + 1
        p[3], p[4], p[5] = 1/rho, 2/rho, rho**2
        p[6], p[7] = (k+1)*rho**2, k*rho**2
        p[8] = self._edge_z()
        p[9] = self._fd_eps()
        # warp-uniform values of the hit test (conic_surface.py:160-161, surface.py:235-245) with the reference's own
        # float64 expressions: z_min - N_EPS, z_max + N_EPS, (r + N_EPS)^2
        rb = float(self.r) + self.N_EPS
        p[11], p[12], p[13] = float(self.z_min) - self.N_EPS, float(self.z_max) + self.N_EPS, rb*rb

    def _record(self):
        rec = self._base_record()
        self._conic_par(rec)
        return rec


class SphericalSurface(ConicSurface):
    """spherical_surface.py"""
    sphere_projection_methods = ["Equidistant", "Orthographic", "Equal-Area", "Stereographic"]

    def __init__(self, r: float, R: float, **kwargs):
        super().__init__(r, R, k=0, **kwargs)

    @property
    def info(self):
        return Surface.info.fget(self) + f", R = {self.R:.5g} mm"

    def sphere_projection(self, p: np.ndarray, projection_method: str = "Equidistant") -> np.ndarray:
        """Projection of 3-D points on the sphere (spherical_surface.py:36-97), evaluated on the device."""
        from . import engine
        return engine.sphere_projection(self, p, projection_method)


class TiltedSurface(Surface):
    """tilted_surface.py"""
    rotational_symmetry = False
    _kind = K_TILTED

    def __init__(self, r: float, normal=None, normal_sph=None, **kwargs):
        super().__init__(r, **kwargs)
        self.parax_roc = None
        self.z_min = self.z_max = self.pos[2]
        if normal is not None:
            self._set_normal(normal)
        elif normal_sph is not None:
            if not isinstance(normal_sph, (list, np.ndarray)):
                raise TypeError("normal_sph needs to be a list or array.")
            theta, phi = np.radians(normal_sph[0]), np.radians(normal_sph[1])
            self._set_normal([np.sin(theta)*np.cos(phi), np.sin(theta)*np.sin(phi), np.cos(theta)])
        else:
            raise RuntimeError("normal or normal_sph parameter needs to be specified.")
        phi = np.arctan2(self.normal[1], self.normal[0])
        R = self.r
        val1 = self.pos[2] + self._values(np.array([R*np.cos(phi)]), np.array([R*np.sin(phi)]))[0]
        val2 = self.pos[2] + self._values(np.array([-R*np.cos(phi)]), np.array([-R*np.sin(phi)]))[0]
        self.z_min, self.z_max = min(val1, val2), max(val1, val2)

    def _set_normal(self, val):
        if not isinstance(val, (list, np.ndarray)):
            raise TypeError("normal needs to be a list or array.")
        val2 = np.asarray_chkfinite(val, dtype=np.float64)/np.linalg.norm(val)
        if not val2[2] > 0:
            raise ValueError("normal[2] needs to be above 0.")
        self.normal = val2

    def _values(self, x, y):
        """tilted_surface.py:61-74"""
        mx = -self.normal[0]/self.normal[2]
        my = -self.normal[1]/self.normal[2]
        return x*mx + y*my

    def flip(self) -> None:
        self.normal = self.normal.copy()
        self.normal[0] *= -1

    def rotate(self, angle: float) -> None:
        self.normal = self.normal.copy()
        self.normal[:2] = _rot(self.normal[0], self.normal[1], np.deg2rad(angle))

    def _record(self):
        rec = self._base_record()
        n = self.normal
        p = rec["par"]
        p[0], p[1], p[2] = float(n[0]), float(n[1]), float(n[2])
        p[3], p[4] = float(-n[0]/n[2]), float(-n[1]/n[2])
        p[9] = self._fd_eps()
        return rec


class FunctionSurface2D(Surface):
    """function_surface_2d.py — user callables are turned into CUDA device functions (userfunc.py)."""
    rotational_symmetry = False
    _1D = False
    _kind = K_FUNC

    def __init__(self, r: float, func: Callable, mask_func: Callable = None, deriv_func: Callable = None,
                 func_args: dict = {}, mask_args: dict = {}, deriv_args: dict = {},
                 z_min: float = None, z_max: float = None, parax_roc: float = None, **kwargs):
        super().__init__(r, **kwargs)
        if not callable(func):
            raise TypeError("func needs to be callable.")
        for name, f in (("mask_func", mask_func), ("deriv_func", deriv_func)):
            if f is not None and not callable(f):
                raise TypeError(f"{name} needs to be callable or None.")
        for name, d in (("func_args", func_args), ("mask_args", mask_args), ("deriv_args", deriv_args)):
            if not isinstance(d, dict):
                raise TypeError(f"{name} needs to be a dict.")
        self._sign = 1
        self._angle = 0
        self.func, self.mask_func, self.deriv_func = func, mask_func, deriv_func
        self._func_args = _copy.deepcopy(func_args)
        self._mask_args = _copy.deepcopy(mask_args)
        self._deriv_args = _copy.deepcopy(deriv_args)
        self._offset = 0
        self._offset = self._values(np.array([0.]), np.array([0.]))[0]
        self.parax_roc = parax_roc
        self._set_zmin_zmax(z_min, z_max)

    def _set_zmin_zmax(self, z_min, z_max):
        """function_surface_2d.py:81-131"""
        if self._1D:
            rn = np.linspace(0, self.r, 10000)
            zn = self._values(rn, np.zeros_like(rn))
            mn = self.mask(rn, np.zeros_like(rn))
            self.z_min, self.z_max = float(zn[mn].min()), float(zn[mn].max())
        else:
            self.z_min, self.z_max = self._find_bounds()
        name = f"{type(self).__name__} {self.get_desc(hex(id(self)))}"
        if z_max is not None and z_min is not None:
            probed = self.z_max - self.z_min
            provided = z_max - z_min
            if probed and provided + self.N_EPS < probed:
                warning(f"{name}: Provided a z-extent of {provided}, but measured range is at least {probed}."
                        f" I will use the measured values for now.")
            else:
                if provided > 1.2*probed:
                    warning(f"{name}: Provided z-range is more than 20% larger than measured z-range")
                z_max_ = self.z_max + self._offset
                z_min_ = self.z_min + self._offset
                if z_max + self.N_EPS < z_max_:
                    warning(f"{name}: Provided z_max={z_max} lower than measured value of {z_max_}."
                            f" Using the measured values for now")
                elif z_min - self.N_EPS > z_min_:
                    warning(f"{name}: Provided z_min={z_min} higher than measured value of {z_min_}."
                            f" Using the measured values for now")
                else:
                    self.z_min, self.z_max = float(z_min - self._offset), float(z_max - self._offset)
        elif z_max is None and z_min is None:
            warning(f"Estimated z-bounds of {name}: [{self._offset+self.z_min:.9g}, "
                    f"{self._offset+self.z_max:.9g}], provide actual values for higher precision.")
        else:
            raise ValueError("z_max and z_min need to be both None or both need a value")

    def _values(self, x, y):
        """function_surface_2d.py:133-156 (host evaluation of the user callable: setup only)"""
        if self._1D:
            r = np.sqrt(x**2 + y**2)
            vals = self.func(r, **self._func_args)
        else:
            x_, y_ = _rot(x, y, -self._angle)
            vals = self.func(x_, self._sign*y_, **self._func_args)
        if not isinstance(vals, np.ndarray):
            raise RuntimeError(f"func must return a np.ndarray, but returns type {type(vals)}.")
        if vals.shape[0] and not isinstance(vals[0], np.float64):
            raise RuntimeError("Elements of return value of func must be of type np.float64")
        return self._sign*(vals - self._offset)

    def mask(self, x, y):
        """function_surface_2d.py:158-191"""
        m = super().mask(x, y)
        if self.mask_func is not None:
            xm, ym = x - self.pos[0], y - self.pos[1]
            if self._1D:
                mf = self.mask_func(np.sqrt(xm**2 + ym**2), **self._mask_args)
            else:
                x_, y_ = _rot(xm, ym, -self._angle)
                mf = self.mask_func(x_, self._sign*y_, **self._mask_args)
            if not isinstance(mf, np.ndarray):
                raise RuntimeError(f"mask_func must return a np.ndarray, but returns type {type(mf)}.")
            m = m & mf
        return m

    def flip(self) -> None:
        """function_surface_2d.py:255-272"""
        self._sign *= -1
        self.parax_roc = self.parax_roc if self.parax_roc is None else -self.parax_roc
        a = self.pos[2] - (self.z_max - self.pos[2])
        b = self.pos[2] - (self.z_min - self.pos[2])
        self.z_min, self.z_max = a, b

    def rotate(self, angle: float) -> None:
        if not self._1D:
            self._angle += np.deg2rad(angle)

    def _func_par(self, rec):
        a = self._angle
        p = rec["par"]
        p[0], p[1] = float(self._sign), float(self._offset)
        if a:
            rec["flags"] |= F_ROTATED
        p[2], p[3] = float(np.cos(-a)), float(np.sin(-a))
        p[4], p[5] = float(np.cos(a)), float(np.sin(a))
        p[8] = self._edge_z() if self.rotational_symmetry else 0.0
        p[9] = self._fd_eps()

    def _record(self):
        rec = self._base_record()
        if self._1D:
            rec["flags"] |= F_1D
        if self.deriv_func is not None:
            rec["flags"] |= F_HAS_DERIV
        if self.mask_func is not None:
            rec["flags"] |= F_HAS_MASK
        self._func_par(rec)
        rec["funcs"] = dict(func=(self.func, self._func_args), mask=(self.mask_func, self._mask_args),
                            deriv=(self.deriv_func, self._deriv_args), one_d=self._1D)
        return rec


class FunctionSurface1D(FunctionSurface2D):
    """function_surface_1d.py"""
    rotational_symmetry = True
    _1D = True


class AsphericSurface(Surface):
    """aspheric_surface.py — conic plus even polynomial; numeric hit finding, analytic radial derivative.
    (The reference derives it from FunctionSurface1D; here it is a closed-form device kind of its own.)"""
    rotational_symmetry = True
    _1D = True
    _kind = K_ASPHERE

    def __init__(self, r: float, R: float, k: float, coeff, **kwargs):
        super().__init__(r, **kwargs)
        for name, v in (("R", R), ("k", k)):
            if not isinstance(v, (int, float)):
                raise TypeError(f"{name} needs to be a number.")
        R, k = float(R), float(k)
        if R == 0 or not np.isfinite(R):
            raise ValueError("R needs to be non-zero and finite. Use planar surface types for planar surfaces.")
        if not isinstance(coeff, (list, np.ndarray)):
            raise TypeError("coeff needs to be a list or array.")
        coeff = np.asarray_chkfinite(coeff, dtype=np.float64)
        if not len(coeff):
            raise ValueError("Empty coeff list. Provide coefficients or use ConicSurface instead.")
        self.R, self.k, self.coeff = R, k, coeff
        self._sign = 1      # flipping negates R and coeff instead (aspheric_surface.py:84-101)
        self._offset = 0
        self._offset = self._values(np.array([0.]), np.array([0.]))[0]
        self.parax_roc = 1/(1/self.R + 2*self.coeff[0])
        rn = np.linspace(0, self.r, 10000)
        zn = self._values(rn, np.zeros_like(rn))
        mn = self.mask(rn, np.zeros_like(rn))
        self.z_min, self.z_max = float(zn[mn].min()), float(zn[mn].max())

    @property
    def _np_coeff(self):
        """aspheric_surface.py:103-112"""
        c = np.zeros(2*len(self.coeff) + 1, dtype=np.float64)
        c[2::2] = self.coeff
        return np.flip(c)

    def _asph(self, r):
        rho, k = 1/self.R, self.k
        z = rho*r**2/(1 + np.sqrt(1 - (k+1)*rho**2*r**2))
        z += np.polyval(self._np_coeff, r)
        return z

    def _values(self, x, y):
        r = np.sqrt(x**2 + y**2)
        return self._sign*(self._asph(r) - self._offset)

    def flip(self) -> None:
        self.R *= -1
        self.coeff = self.coeff*-1
        self.parax_roc *= -1
        a = self.pos[2] - (self.z_max - self.pos[2])
        b = self.pos[2] + (self.pos[2] - self.z_min)
        self.z_min, self.z_max = a, b

    def rotate(self, angle: float) -> None:
        pass

    @property
    def info(self):
        return super().info + f", R = {self.R:.5g} mm, k = {self.k:.5g}\ncoeff = {self.coeff}"

    def _record(self):
        rec = self._base_record()
        rec["flags"] |= F_1D | F_HAS_DERIV
        k, rho = self.k, 1/self.R
        p = rec["par"]
        p[0], p[1], p[2] = k, rho, k + 1
        p[5], p[6] = rho**2, (k+1)*rho**2
        p[8] = self._edge_z()
        p[9] = self._fd_eps()
        p[10] = float(self._offset)
        npc = self._np_coeff
        der = np.polyder(npc)
        rec["aux"] = np.concatenate((npc, der)).astype(np.float64)
        rec["aux_n0"], rec["aux_n1"] = len(npc), len(der)
        return rec


class DataSurface2D(Surface):
    """data_surface_2d.py — quartic FITPACK spline of a height grid.  The fit is scene setup (scipy on the
    host, like the reference); evaluation on the ray path is a degree-4 de Boor on the device from the
    extracted knots/coefficients (SURVEY.md hard part 4)."""
    rotational_symmetry = False
    _1D = False
    _kind = K_DATA

    def __init__(self, r: float, data, parax_roc: float = None, **kwargs):
        import scipy.interpolate
        super().__init__(r, **kwargs)
        self._sign = 1
        self._angle = 0
        self._interp, self._offset = None, 0.
        self.parax_roc = parax_roc
        if not isinstance(data, (np.ndarray, list)):
            raise TypeError("data needs to be an array or list.")
        Z = np.array(np.asarray_chkfinite(data, dtype=np.float64))
        name = f"{type(self).__name__} {self.get_desc(hex(id(self)))}"
        nx = Z.shape[0]
        if nx < 50:
            raise ValueError("For a good surface representation 'data' should have at least 50 values per dimension")
        if nx < 200:
            warning(f"{name}: At least 200 values per dimension are advised for a 'data' matrix, but got {nx} values.")
        if self._1D:
            if Z.ndim != 1:
                raise ValueError("data array needs to have exactly one dimension.")
            Z -= Z[0]
            r0 = np.linspace(0, self.r, Z.shape[0], dtype=np.float64)
            r2 = np.concatenate((-np.flip(r0[1:]), r0))
            z2 = np.concatenate((np.flip(Z[1:]), Z))
            self._interp = scipy.interpolate.InterpolatedUnivariateSpline(r2, z2, k=4)
            self._offset = self._call(0, 0)
            rn = np.linspace(0, self.r, 10000)
            zn = self._values(rn, np.zeros_like(rn))
            self.z_min, self.z_max = float(np.min(zn)), float(np.max(zn))
            z_range0 = np.ptp(Z)
        else:
            if Z.ndim != 2:
                raise ValueError("data array needs to have exactly two dimensions.")
            ny, nx = Z.shape
            if nx != ny:
                raise ValueError("Array 'data' needs to be of square shape.")
            if nx % 2:
                Z -= np.array([Z[ny//2, nx//2], Z[ny//2+1, nx//2], Z[ny//2, nx//2+1], Z[ny//2+1, nx//2+1]]).mean()
            else:
                Z -= Z[ny//2, nx//2]
            xy = np.linspace(-self.r, self.r, nx)
            self._interp = scipy.interpolate.RectBivariateSpline(xy, xy, Z, kx=4, ky=4)
            self._offset = self._call(0, 0)
            self.z_min, self.z_max = self._find_bounds()
            X, Y = np.meshgrid(xy, xy)
            M = self.mask(X.ravel(), Y.ravel()).reshape(X.shape)
            z_range0 = np.max(Z[M]) - np.min(Z[M])
        z_range1 = self.z_max - self.z_min
        if np.abs(z_range0 - z_range1) > self.N_EPS:
            z_change = (z_range1 - z_range0)/z_range0
            warning(f"{name}: Due to interpolation the z_range of the surface has increased from "
                    f"{z_range0:.9g} to {z_range1:.9g}, a change of {z_change*100:.5g}%.")

    def _call(self, x, y, **kw):
        if self._1D:
            return self._interp(np.hypot(x, y), **kw)
        return self._interp(x, y, grid=False, **kw)

    def _values(self, x, y):
        """data_surface_2d.py:140-153"""
        x_, y_ = _rot(x, y, -self._angle) if not self.rotational_symmetry else (x, y)
        return self._sign*(self._call(x_, self._sign*y_) - self._offset)

    def flip(self) -> None:
        self._sign *= -1
        self.parax_roc = self.parax_roc if self.parax_roc is None else -self.parax_roc
        a = self.pos[2] - (self.z_max - self.pos[2])
        b = self.pos[2] - (self.z_min - self.pos[2])
        self.z_min, self.z_max = a, b

    def rotate(self, angle: float) -> None:
        if not self.rotational_symmetry:
            self._angle += np.deg2rad(angle)

    def _record(self):
        rec = self._base_record()
        a = self._angle
        p = rec["par"]
        p[0], p[1] = float(self._sign), float(self._offset)
        if a:
            rec["flags"] |= F_ROTATED
        p[2], p[3] = float(np.cos(-a)), float(np.sin(-a))
        p[4], p[5] = float(np.cos(a)), float(np.sin(a))
        p[8] = self._edge_z() if self.rotational_symmetry else 0.0
        p[9] = self._fd_eps()
        if self._1D:
            rec["flags"] |= F_1D
            t, c, k = self._interp._eval_args
            assert k == 4
            rec["aux"] = np.concatenate((np.asarray(t, dtype=np.float64), np.asarray(c, dtype=np.float64)[:len(t)-5]))
            rec["aux_n0"], rec["aux_n1"] = len(t), 0
        else:
            tx, ty, c = self._interp.tck
            assert tuple(self._interp.degrees) == (4, 4)
            rec["aux"] = np.concatenate((np.asarray(tx, dtype=np.float64), np.asarray(ty, dtype=np.float64),
                                         np.asarray(c, dtype=np.float64)))
            rec["aux_n0"], rec["aux_n1"] = len(tx), len(ty)
        return rec


class DataSurface1D(DataSurface2D):
    """data_surface_1d.py"""
    rotational_symmetry = True
    _1D = True
