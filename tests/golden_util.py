"""helpers shared by the oracle-pinning tests and the GPU parity tests"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"
PROJ = ["Equidistant", "Orthographic", "Equal-Area", "Stereographic"]


def load(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


def bundle(g):
    return g["p0"], g["s0"], g.get("pol0"), g["w0"], g["wl"], g.get("hurb_z")


def maxrel(a, b):
    """largest |a-b| / max(|b|, tiny) over finite entries; NaN patterns must coincide"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return np.inf
    m = ~na
    if not m.any():
        return 0.0
    d = np.abs(a[m] - b[m])
    return float(np.max(d/np.maximum(np.abs(b[m]), 1e-300)*(d > 0)))


def vecrel(a, b):
    """largest |a-b| relative to the length of the reference 3-vector (last axis); NaN patterns must coincide.
    Element-wise relative errors are meaningless for vector components that pass through zero."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return np.inf
    d = np.where(na, 0.0, np.abs(a - np.where(nb, 0.0, b)))
    nrm = np.sqrt(np.nansum(np.where(nb, 0.0, b)**2, axis=-1, keepdims=True))
    return float(np.max(d/np.maximum(nrm, 1e-300)*(d > 0)))


# Per-scene weight tolerance.  The reference evaluates a Gaussian TransmissionSpectrum on its raw float32
# wavelengths (spectrum.py:113 + NEP 50), i.e. with numpy's float32 exp; the engine rounds a float64 exp to
# float32.  Both are float32-accurate, they differ by at most ~2 float32 ulp (parity notes).
W_RTOL = {"zoo_analytic": 3e-7}
