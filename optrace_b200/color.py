"""Colour data used ON the hot path: CIE observer tables (detector binning), illuminant tables
(preset source spectra), sRGB primaries (image-source wavelength sampling).

Everything here is table construction at scene build time; per-ray use of these tables
(observer interpolation, inverse-CDF lookups) happens on the device.
References: optrace/tracer/color/observers.py, illuminants.py, tools.py, srgb.py:24-46, 456-565.
The full colour-science stack of the reference (sRGB/Luv conversion of finished images) is
out of scope.
"""
import pathlib

import numpy as np

from .options import global_options as go

_tab = np.load(pathlib.Path(__file__).resolve().parent / "data" / "cie_tables.npz")
OBSERVERS = np.ascontiguousarray(_tab["observers"])        # (471, 4): wl, x, y, z  (360..830 nm, 1 nm)
ILLUMINANTS = np.ascontiguousarray(_tab["illuminants"])
ILLUMINANT_NAMES = [str(s) for s in _tab["illuminant_names"]]

WL_MIN0, WL_MAX0 = 380., 780.

SRGB_R_PRIMARY_POWER_FACTOR = 0.885651229244
SRGB_G_PRIMARY_POWER_FACTOR = 1.000000000000
SRGB_B_PRIMARY_POWER_FACTOR = 0.775993481741


def wavelengths(N: int) -> np.ndarray:
    return np.linspace(*go.wavelength_range, N)


def x_observer(wl):
    return np.interp(wl, OBSERVERS[:, 0], OBSERVERS[:, 1], left=0, right=0)


def y_observer(wl):
    return np.interp(wl, OBSERVERS[:, 0], OBSERVERS[:, 2], left=0, right=0)


def z_observer(wl):
    return np.interp(wl, OBSERVERS[:, 0], OBSERVERS[:, 3], left=0, right=0)


def illuminant(name: str):
    """returns f(wl) interpolating the CIE table of illuminant `name` (illuminants.py)"""
    if name == "E":
        return lambda wl: np.full_like(wl, 100.0, dtype=np.float64)
    col = ILLUMINANT_NAMES.index(name)

    def f(wl):
        return np.interp(wl, ILLUMINANTS[:, 0], ILLUMINANTS[:, col], left=0, right=0)
    f.__name__ = f"{name.lower().replace('-', '_')}_illuminant"
    return f


d65_illuminant = illuminant("D65")


def blackbody(wl: np.ndarray, T: float = 6504.) -> np.ndarray:
    """Planck curve, W/(sr m^3) (tools.py:25-43)"""
    import scipy.constants
    c, h, k_B = scipy.constants.c, scipy.constants.h, scipy.constants.k
    wlm = 1e-9*wl
    return 2*h*c**2/wlm**5/(np.exp(h*c/(wlm*k_B*T)) - 1)


def normalized_blackbody(wl: np.ndarray, T: float = 6504.) -> np.ndarray:
    """tools.py:45-59"""
    l_w = 2897.771955*1e3/T
    p_w, p_l, p_r = blackbody(np.array([l_w, *go.wavelength_range]), T)
    p_max = p_w if go.wavelength_range[0] <= l_w <= go.wavelength_range[1] else max(p_l, p_r)
    return blackbody(wl, T)/p_max


def srgb_to_srgb_linear(rgb: np.ndarray) -> np.ndarray:
    """srgb.py:29-46"""
    a = 0.055
    below = np.abs(rgb) <= 0.04045
    lin = np.sign(rgb)*(1/(1 + a)*(np.abs(rgb) + a))**2.4
    lin[below] = 1/12.92*rgb[below]
    return lin


def power_from_srgb_linear(rgbl: np.ndarray) -> np.ndarray:
    """srgb.py:556-565"""
    return (SRGB_R_PRIMARY_POWER_FACTOR*rgbl[:, :, 0] + SRGB_G_PRIMARY_POWER_FACTOR*rgbl[:, :, 1]
            + SRGB_B_PRIMARY_POWER_FACTOR*rgbl[:, :, 2])


def _gauss(x, mu, sig):
    return 1/(sig*np.sqrt(2*np.pi))*np.exp(-0.5/sig**2*(x - mu)**2)


def _clip_vis(wl, v):
    v[~((wl >= WL_MIN0) & (wl <= WL_MAX0))] = 0
    return v


def srgb_r_primary(wl):
    """srgb.py:469-481"""
    rs = 0.951190393
    return _clip_vis(wl, 75.1660756583*rs*(_gauss(wl, 639.854491, 30.0) + 0.0500907584*_gauss(wl, 418.905848, 80.6220465)))


def srgb_g_primary(wl):
    """srgb.py:484-495"""
    return _clip_vis(wl, 83.4999222966*1*_gauss(wl, 539.13108974, 33.31164968))


def srgb_b_primary(wl):
    """srgb.py:498-509"""
    bs = 1.16364585503
    return _clip_vis(wl, 47.99521746361*bs*(_gauss(wl, 454.833119, 20.1460206) + 0.184484176*_gauss(wl, 459.658190, 71.0927568)))


def srgb_primary_cdfs():
    """(wl[5000], F_r, F_g, F_b): cumulative trapezoid tables of the three primaries
    (srgb.py:528, 549-551 with random.py:143-157)."""
    import scipy.integrate
    wl = wavelengths(5000)
    return (wl,) + tuple(scipy.integrate.cumulative_trapezoid(f(wl), initial=0)
                         for f in (srgb_r_primary, srgb_g_primary, srgb_b_primary))
