"""Spectra and refraction indices of the drop-in API (host side).

Constructor signatures follow optrace/tracer/spectrum/spectrum.py, light_spectrum.py,
transmission_spectrum.py and optrace/tracer/refraction_index.py.  Per-ray evaluation
(n(lambda) at every lens, filter transmission, wavelength sampling) happens on the device;
the classes here flatten themselves into OtbMedium / OtbFilter records and into inverse-CDF
tables for the on-device ray generator.  Host evaluation of `Spectrum.__call__` is limited to
table construction at scene build time (O(1e4) points, the reference does the same per
create_rays call: light_spectrum.py:128-133).
"""
from __future__ import annotations

import copy as _copy
from typing import Callable

import numpy as np

from . import _state

from .options import global_options as go

# OtbMediumModel
N_MODELS = {"Constant": 0, "Abbe": 1, "Cauchy": 2, "Conrady": 3, "Sellmeier1": 4, "Sellmeier2": 5,
            "Sellmeier3": 6, "Sellmeier4": 7, "Sellmeier5": 8, "Schott": 9, "Herzberger": 10,
            "Handbook of Optics 1": 11, "Handbook of Optics 2": 12, "Extended": 13, "Extended2": 14,
            "Extended3": 15, "Data": 16, "Function": 17}
# OtbSpectrumType
T_TYPES = {"Constant": 0, "Data": 1, "Rectangle": 2, "Gaussian": 3, "Function": 4}

# Fraunhofer lines F, d, C in nm (presets/spectral_lines.py)
LINES_FDC = [486.1327, 587.5618, 656.272]


def wavelengths(N: int) -> np.ndarray:
    """color/tools.py:14-22"""
    return np.linspace(*go.wavelength_range, N)


class _Described:
    def __init__(self, desc: str = "", long_desc: str = ""):
        self.desc, self.long_desc = desc, long_desc

    def __setattr__(self, key, val):
        object.__setattr__(self, key, val)
        _state.EPOCH[0] += 1        # scene epoch: see _state.py

    def copy(self):
        return _copy.deepcopy(self)

    def get_desc(self, fallback: str = "") -> str:
        return self.desc if self.desc != "" else fallback

    def get_long_desc(self, fallback: str = "") -> str:
        return self.long_desc if self.long_desc != "" else self.get_desc(fallback)


class Spectrum(_Described):
    """spectrum/spectrum.py"""
    spectrum_types = ["Monochromatic", "Constant", "Data", "Lines", "Rectangle", "Gaussian", "Function"]
    unit = ""
    quantity = ""

    def __init__(self, spectrum_type: str = "Gaussian", val: float = 1., lines=None, line_vals=None,
                 wl: float = 550., wl0: float = 400., wl1: float = 600., wls=None, vals=None,
                 func: Callable = None, mu: float = 550., sig: float = 50., unit: str = None,
                 quantity: str = None, func_args: dict = {}, **kwargs):
        super().__init__(**kwargs)
        if spectrum_type not in self.spectrum_types:
            raise ValueError(f"Invalid spectrum_type '{spectrum_type}', must be one of {self.spectrum_types}.")
        self.spectrum_type = spectrum_type
        # the reference stores lines / line_vals as float32 arrays (spectrum.py:146-148); the Abbe fit and the
        # line sampling tables inherit that rounding
        self.lines = None if lines is None else np.asarray_chkfinite(lines, dtype=np.float32)
        self.line_vals = None if line_vals is None else np.asarray_chkfinite(line_vals, dtype=np.float32)
        self.func_args = _copy.deepcopy(func_args)
        if func is not None and not callable(func):
            raise TypeError("func needs to be callable.")
        self.func = func
        self.wl, self.wl0, self.wl1 = float(wl), float(wl0), float(wl1)
        self.val, self.mu, self.sig = float(val), float(mu), float(sig)
        self._wls = None if wls is None else np.asarray_chkfinite(wls, dtype=np.float64)
        self._vals = None if vals is None else np.asarray_chkfinite(vals, dtype=np.float64)
        self.unit = unit if unit is not None else self.unit
        self.quantity = quantity if quantity is not None else self.quantity
        self._validate()

    def _validate(self):
        r0, r1 = go.wavelength_range
        if self.val < 0:
            raise ValueError("val needs to be non-negative.")
        if self.sig <= 0:
            raise ValueError("sig needs to be above 0.")
        for name in ("wl", "wl0", "wl1", "mu"):
            v = getattr(self, name)
            if not (r0 <= v <= r1):
                raise ValueError(f"{name} needs to be inside the wavelength range [{r0}, {r1}].")
        if self.wl1 <= self.wl0 and self.spectrum_type == "Rectangle":
            raise ValueError("wl1 needs to be above wl0.")
        if self.lines is not None and len(self.lines):
            if min(self.lines) < r0 or max(self.lines) > r1:
                raise ValueError("lines need to be inside the wavelength range.")
        if self.line_vals is not None and len(self.line_vals) and min(self.line_vals) <= 0:
            raise ValueError("line_vals need to be positive.")
        if self._vals is not None and np.min(self._vals) < 0:
            raise ValueError("vals needs to be non-negative.")

    def is_continuous(self) -> bool:
        return self.spectrum_type not in ["Lines", "Monochromatic"]

    def __call__(self, wl):
        """spectrum.py:81-119 — host evaluation for table construction at scene build time."""
        if not self.is_continuous():
            raise RuntimeError(f"Can't call discontinuous spectrum_type '{self.spectrum_type}'")
        wl_ = np.asarray_chkfinite(wl, dtype=np.float64)
        t = self.spectrum_type
        if t == "Constant":
            return np.broadcast_to(self.val, wl_.shape)
        if t == "Data":
            return np.interp(wl_, self._wls, self._vals, left=0, right=0)
        if t == "Rectangle":
            res = np.zeros_like(wl_, dtype=np.float64)
            res[(self.wl0 <= wl_) & (wl_ <= self.wl1)] = self.val
            return res
        if t == "Gaussian":
            return self.val*np.exp(-(wl_ - self.mu)**2/(2*self.sig**2))
        if t == "Function":
            if not callable(self.func):
                raise TypeError("Spectrum.func needs to be callable.")
            return self.func(wl_, **self.func_args)
        raise ValueError(f"Invalid spectrum_type {t}.")

    def get_desc(self, fallback: str = None) -> str:
        fb = str(self.val) if self.spectrum_type == "Constant" else self.spectrum_type
        return super().get_desc(fb)


class TransmissionSpectrum(Spectrum):
    """spectrum/transmission_spectrum.py"""
    spectrum_types = ["Constant", "Data", "Rectangle", "Gaussian", "Function"]
    quantity = "Transmission T"

    def __init__(self, spectrum_type: str = "Gaussian", inverse: bool = False, **sargs):
        if not isinstance(inverse, bool):
            raise TypeError("inverse needs to be bool.")
        self.inverse = inverse
        super().__init__(spectrum_type, **sargs)
        if self.val > 1:
            raise ValueError("val needs to be at most 1.")
        if self._vals is not None and np.max(self._vals) > 1:
            raise ValueError("all elements in vals need to be in range [0, 1].")
        if callable(self.func):
            T = self.func(wavelengths(1000), **self.func_args)
            if np.any(T > 1) or np.any(T < 0):
                raise RuntimeError("Function func needs to return values in range [0, 1] over the visible range.")

    def __call__(self, wl):
        v = super().__call__(wl)
        return v if not self.inverse else 1.0 - v

    def _record(self) -> dict:
        """dict form of OtbFilter"""
        t = self.spectrum_type
        rec = dict(type=T_TYPES[t], inverse=int(self.inverse), func=None, aux=None, c=[0.0]*4)
        if t == "Constant":
            rec["c"][0] = self.val
        elif t == "Rectangle":
            rec["c"][:3] = [self.wl0, self.wl1, self.val]
        elif t == "Gaussian":
            rec["c"][:4] = [self.val, self.mu, self.sig, 2*self.sig**2]
        elif t == "Data":
            rec["aux"] = np.concatenate((self._wls, self._vals))
        elif t == "Function":
            rec["func"] = (self.func, self.func_args)
        return rec


class RefractionIndex(Spectrum):
    """refraction_index.py"""
    coeff_count = {"Cauchy": 4, "Conrady": 3, "Sellmeier1": 6, "Sellmeier2": 5, "Sellmeier3": 8,
                   "Sellmeier4": 5, "Sellmeier5": 10, "Herzberger": 6, "Extended": 8, "Extended2": 8,
                   "Handbook of Optics 1": 4, "Handbook of Optics 2": 4, "Schott": 6, "Extended3": 9}
    n_types = ["Abbe", "Cauchy", "Conrady", "Constant", "Data", "Extended", "Extended2", "Extended3", "Function",
               "Handbook of Optics 1", "Handbook of Optics 2", "Sellmeier1", "Sellmeier2", "Sellmeier3",
               "Sellmeier4", "Sellmeier5", "Herzberger", "Schott"]
    spectrum_types = n_types
    quantity = "Refraction Index n"

    def __init__(self, n_type: str = "Constant", n: float = 1.0, coeff: list = None, lines=None,
                 V: float = None, **kwargs):
        if n_type not in self.n_types:
            raise ValueError(f"Invalid n_type '{n_type}', must be one of {self.n_types}.")
        if coeff is not None:
            if not isinstance(coeff, list):
                raise TypeError("coeff needs to be a list.")
            cnt = self.coeff_count.get(n_type)
            if cnt is not None and len(coeff) != cnt:
                raise ValueError(f"coeff needs to be a list with exactly {cnt} numeric coefficients for mode "
                                 f"{n_type}, but got {len(coeff)}.")
            coeff = list(coeff)
        self.coeff = coeff
        if V is not None:
            if not isinstance(V, (int, float)):
                raise TypeError("V needs to be a number.")
            if V <= 0 or not np.isfinite(V):
                raise ValueError("V needs to be above 0 and finite.")
        self.V = V
        if not isinstance(n, (int, float)):
            raise TypeError("n needs to be a number.")
        if not np.isfinite(n) or n < 1:
            raise ValueError("n needs to be at least 1.")
        lines = lines if lines is not None else LINES_FDC
        if len(lines) != 3:
            raise ValueError("Property 'lines' for n_type='Abbe' needs to have exactly 3 elements")
        if not lines[0] < lines[1] < lines[2]:
            raise ValueError("The values of property 'lines' need to be ascending.")
        super().__init__(n_type, val=n, lines=lines, **kwargs)
        if self._vals is not None and np.min(self._vals) < 1:
            raise ValueError("all vals values needs to be at least 1.")
        if callable(self.func):
            nv = self.func(wavelengths(1000), **self.func_args)
            if np.min(nv) < 1:
                raise ValueError("Function func needs to output values >= 1 over the whole visible range.")

    def _validate(self):
        pass  # range checks of Spectrum do not apply to the index models

    def _abbe_AB(self):
        """refraction_index.py:86-100: Cauchy/Herzberger compromise curve through (nc, V)."""
        l = 1e-3*np.array(self.lines)
        nc = self.val
        d = 0.014
        B = 1/self.V*(nc - 1)/(1/(l[0]**2 - d) - 1/(l[2]**2 - d))
        A = nc - B/(l[1]**2 - d)
        return float(A), float(B), d

    def _record(self) -> dict:
        """dict form of OtbMedium"""
        t = self.spectrum_type
        rec = dict(model=N_MODELS[t], func=None, aux=None, c=[0.0]*12)
        if t == "Constant":
            rec["c"][0] = self.val
        elif t == "Abbe":
            if self.V is None:
                raise TypeError("Abbe number V needs to be provided for n_type='Abbe'.")
            rec["c"][:3] = self._abbe_AB()
        elif t == "Data":
            rec["aux"] = np.concatenate((self._wls, self._vals))
        elif t == "Function":
            rec["func"] = (self.func, self.func_args)
        else:
            if self.coeff is None:
                raise TypeError(f"coefficient variable 'coeff' needs to be provided for n_type='{t}'.")
            rec["c"][:len(self.coeff)] = [float(v) for v in self.coeff]
        return rec

    def __call__(self, wl):
        """n(lambda) evaluated by the CUDA engine (refraction_index.py:62-169); raises like the reference for n < 1."""
        from . import engine
        wl_ = np.atleast_1d(np.asarray_chkfinite(wl, dtype=np.float64))
        if self.spectrum_type == "Data":
            if wl_.min() < self._wls[0] or wl_.max() > self._wls[-1]:
                raise RuntimeError(f"Wavelength range [{wl_.min():.5g}, {wl_.max():.5g}] larger than data range"
                                   f" [{self._wls[0]}, {self._wls[-1]}] for this material.")
        ns = engine.medium_eval(self, wl_)
        i = np.argmin(ns)
        if ns.flat[i] < 1:
            raise RuntimeError(f"Refraction index below 1 with value {ns.flat[i]:.4g} at {wl_.flat[i]:.4g}nm.")
        return ns.reshape(np.shape(wl)) if np.ndim(wl) else ns

    def __eq__(self, other) -> bool:
        if type(self) is not type(other):
            return False
        if self is other:
            return True
        a, b = self._record(), other._record()
        if a["model"] != b["model"] or a["c"] != b["c"]:
            return False
        if (a["aux"] is None) != (b["aux"] is None):
            return False
        if a["aux"] is not None and not np.array_equal(a["aux"], b["aux"]):
            return False
        if a["func"] is not None and (a["func"][0] is not b["func"][0] or a["func"][1] != b["func"][1]):
            return False
        return True

    def __ne__(self, other) -> bool:
        return not self.__eq__(other)

    __hash__ = object.__hash__

    def abbe_number(self, lines: list = None) -> float:
        lines = lines if lines is not None else self.lines
        ns, nc, nl = tuple(self(np.array(lines, dtype=np.float64)))
        return float((nc - 1)/(ns - nl) if ns != nl else np.inf)

    def is_dispersive(self) -> bool:
        return bool(np.isfinite(self.abbe_number()))


class LightSpectrum(Spectrum):
    """spectrum/light_spectrum.py — sampling happens on the device from the tables built here."""
    spectrum_types = [*Spectrum.spectrum_types, "Blackbody", "Histogram"]

    def __init__(self, spectrum_type: str = "Blackbody", T: float = 5500, **sargs):
        if not isinstance(T, (int, float)) or T <= 0:
            raise ValueError("T needs to be a positive number.")
        self.T = float(T)
        line_spec = spectrum_type in ["Monochromatic", "Lines"]
        super().__init__(spectrum_type, unit="W" if line_spec else "W/nm",
                         quantity="Spectral Power" if line_spec else "Spectral Power Density", **sargs)

    def __call__(self, wl):
        """light_spectrum.py:140-165"""
        if self.spectrum_type == "Blackbody":
            from .color import normalized_blackbody
            wl_ = np.asarray_chkfinite(wl, dtype=np.float64)
            return self.val*normalized_blackbody(wl_, T=self.T)
        if self.spectrum_type == "Histogram":
            wl_ = np.asarray_chkfinite(wl, dtype=np.float64)
            if self._wls is None or self._vals is None:
                raise RuntimeError("Histogram spectrum without data.")
            ind = np.clip(np.searchsorted(self._wls, wl_, side="right") - 1, 0, self._vals.shape[0] - 1)
            res = self._vals[ind].astype(np.float64)
            res[(wl_ < self._wls[0]) | (wl_ > self._wls[-1])] = 0
            return res
        return super().__call__(wl)

    def _sampling_record(self) -> dict:
        """Wavelength sampling description for the device generator (OtbWavelengthMode + table).
        Mirrors LightSpectrum.random_wavelengths (light_spectrum.py:81-138) and the CDF construction of
        random.inverse_transform_sampling (random.py:113-159)."""
        import scipy.integrate
        import scipy.special
        t = self.spectrum_type
        if t == "Monochromatic":
            return dict(mode=0, wl=[self.wl, 0, 0, 0], tab=None)
        if t in ("Constant", "Rectangle"):
            wl0 = go.wavelength_range[0] if t == "Constant" else self.wl0
            wl1 = go.wavelength_range[1] if t == "Constant" else self.wl1
            return dict(mode=1, wl=[wl0, wl1, 0, 0], tab=None)
        if t == "Lines":
            x, f = self.lines, self.line_vals      # float32, cumulated in float32 like random.py:133
            x, f = x[f > 0], f[f > 0]
            return dict(mode=2, wl=[0, 0, 0, 0], tab=(x.astype(np.float64), np.cumsum(f).astype(np.float64)))
        if t == "Gaussian":
            Xl = (1 + scipy.special.erf((go.wavelength_range[0] - self.mu)/(np.sqrt(2)*self.sig)))/2
            Xr = (1 + scipy.special.erf((go.wavelength_range[1] - self.mu)/(np.sqrt(2)*self.sig)))/2
            return dict(mode=4, wl=[self.mu, self.sig, float(Xl), float(Xr)], tab=None)
        if t == "Data":
            x, f = self._wls, self._vals
        else:  # Blackbody, Function, Histogram
            x = wavelengths(4000 if t == "Blackbody" else 10000)
            f = np.asarray(self(x), dtype=np.float64)
        if not f.sum():
            raise RuntimeError("Cumulated probability is zero.")
        if f.min() < 0:
            raise RuntimeError("Got negative value in pdf.")
        F = scipy.integrate.cumulative_trapezoid(f, initial=0)
        return dict(mode=3, wl=[0, 0, 0, 0], tab=(np.asarray(x, dtype=np.float64), F))

    @staticmethod
    def render(wl: np.ndarray, w: np.ndarray, **kwargs) -> "LightSpectrum":
        """light_spectrum.py:40-79: histogram spectrum from ray wavelengths and powers."""
        spec = LightSpectrum("Histogram", **kwargs)
        N = max(51, np.sqrt(np.count_nonzero(w))/2)
        N = 1 + 2*(int(N)//2)
        if not wl.shape[0]:
            spec._wls = wavelengths(N + 1)
            spec._vals = np.zeros(N, dtype=np.float64)
        else:
            wl0, wl1 = wl.min(), wl.max()
            if np.abs(wl0 - wl1) < 1:
                wl0, wl1 = max(wl0 - 1, go.wavelength_range[0]), min(wl0 + 1, go.wavelength_range[1])
            vals, wls = np.histogram(wl, bins=N, weights=w, range=[wl0, wl1])
            # Spectrum.__setattr__ (spectrum.py:184-187) stores both as float64 arrays BEFORE the scaling
            spec._wls = np.asarray(wls, dtype=np.float64)
            spec._vals = np.asarray(vals, dtype=np.float64)*(1/(spec._wls[1] - spec._wls[0]))
        return spec

    @staticmethod
    def _render_device(lib, wl, w, positive_only: bool, **kwargs) -> "LightSpectrum":
        """LightSpectrum.render on device-resident float32 wavelengths / weights (engine.spectrum_histogram):
        same bin count, range, float32 edges and index rule as the host path above; bin sums accumulate in float64
        on the device (the reference accumulates 65536-ray blocks in float32, np.histogram with float32 weights)."""
        from . import engine
        spec = LightSpectrum("Histogram", **kwargs)
        vals, edges = engine.spectrum_histogram(lib, wl, w, positive_only, go.wavelength_range)
        if vals is None:
            N = edges
            spec._wls = wavelengths(N + 1)
            spec._vals = np.zeros(N, dtype=np.float64)
        else:
            spec._wls = np.asarray(edges, dtype=np.float64)
            spec._vals = vals.astype(np.float32).astype(np.float64)*(1/(spec._wls[1] - spec._wls[0]))
        return spec
