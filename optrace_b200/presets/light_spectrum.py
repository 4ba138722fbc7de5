"""Preset light spectra: CIE standard illuminants as "Function" spectra interpolating the CIE tables
(reference: presets/light_spectrum.py, color/illuminants.py)."""
from ..media import LightSpectrum
from .. import color as _color
from . import spectral_lines as _L


def _ill(name, long_desc=None):
    return LightSpectrum("Function", func=_color.illuminant(name), desc=name, long_desc=long_desc or f"Illuminant {name}")


a, c, e = _ill("A"), _ill("C"), _ill("E")
d50, d55, d65, d75 = _ill("D50"), _ill("D55"), _ill("D65"), _ill("D75")
f2, f7, f11 = _ill("F2"), _ill("F7"), _ill("F11")
led_b1, led_b2, led_b3, led_b4, led_b5 = (_ill(f"LED-B{i}") for i in range(1, 6))
led_bh1, led_rgb1, led_v1, led_v2 = _ill("LED-BH1"), _ill("LED-RGB1"), _ill("LED-V1"), _ill("LED-V2")
standard = [a, c, d50, d55, d65, d75, e, f2, f7, f11, led_b1, led_b2, led_b3, led_b4, led_b5,
            led_bh1, led_rgb1, led_v1, led_v2]

FDC = LightSpectrum("Lines", lines=_L.FDC, line_vals=[1, 1, 1], desc="Lines FDC")
FdC = LightSpectrum("Lines", lines=_L.FdC, line_vals=[1, 1, 1], desc="Lines FdC")
FeC = LightSpectrum("Lines", lines=_L.FeC, line_vals=[1, 1, 1], desc="Lines FeC")
F_eC_ = LightSpectrum("Lines", lines=_L.F_eC_, line_vals=[1, 1, 1], desc="Lines F'eC'")
