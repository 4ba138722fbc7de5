"""TEST INFRASTRUCTURE — CPU restatement of LightSpectrum.render (optrace/tracer/spectrum/light_spectrum.py:40-79):
the weighted wavelength histogram behind Raytracer.detector_spectrum / source_spectrum (raytracer.py:1100-1132,
1311-1329).  Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this module.

Pinned: tests/test_oracle_golden.py compares it with spectra the reference itself rendered from the fixture rays
(tests/golden/*.npz keys spec_*; generator tools/gen_golden.py)."""
import numpy as np

WAVELENGTH_RANGE = (380.0, 780.0)      # global_options.wavelength_range (global_options.py)


def render(wl: np.ndarray, w: np.ndarray, wavelength_range=WAVELENGTH_RANGE):
    """(vals, wls): histogram values in W/nm and the bin edges, dtypes as numpy produces them for float32 input"""
    # light_spectrum.py:57-58: at least 51 bins, sqrt(N)/2 above that, odd
    N = max(51, np.sqrt(np.count_nonzero(w))/2)
    N = 1 + 2*(int(N)//2)
    if not wl.shape[0]:                                         # :61-63
        return np.zeros(N, dtype=np.float64), np.linspace(*wavelength_range, N + 1)
    wl0, wl1 = wl.min(), wl.max()                               # :67
    if np.abs(wl0 - wl1) < 1:                                   # :70-71
        wl0, wl1 = max(wl0 - 1, wavelength_range[0]), min(wl0 + 1, wavelength_range[1])
    vals, wls = np.histogram(wl, bins=N, weights=w, range=[wl0, wl1])       # :73
    # assignment to spec._vals / spec._wls converts to float64 (Spectrum.__setattr__, spectrum.py:184-187)
    vals, wls = np.asarray(vals, dtype=np.float64), np.asarray(wls, dtype=np.float64)
    vals = vals*(1/(wls[1] - wls[0]))                           # :74
    return vals, wls
