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
    """the injected bundle of a fixture; the large fixtures keep only s0 and rebuild the rest from section 0"""
    if "p0" not in g:
        pol0 = np.ascontiguousarray(g["pol_list"][:, 0]) if "pol_list" in g else None
        return np.ascontiguousarray(g["p_list"][:, 0]), g["s0"], pol0, np.ascontiguousarray(g["w_list"][:, 0]), g["wl"], g.get("hurb_z")
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


def f32_flips(a, b, max_frac=1e-4):
    """Quantities the reference STORES as float32 (weights, polarisation; raytracer.py:309-312, 828, 876) on scenes
    whose float64 arithmetic is not bit-reproducible (transcendental functions: cos / atan2 / splines agree with
    numpy to an ulp, not bit for bit): a float64 value within 1e-12 of a float32 rounding boundary may round to
    the NEIGHBOURING float32 (relative step 6e-8 — no 1e-9 comparison of the stored value can hold for it).
    Returns a copy of `a` with those entries (differing from b by exactly one float32 ulp) replaced by b's, after
    asserting that they are the rare exception (at most max_frac of the entries); everything else is then compared
    at the usual 1e-9."""
    a32, b32 = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        one_ulp = (a32 != b32) & ((np.nextafter(a32, b32) == b32))
    frac = np.count_nonzero(one_ulp)/max(1, a32.size)
    assert frac <= max_frac, f"{np.count_nonzero(one_ulp)} of {a32.size} float32 entries differ by one ulp"
    out = np.array(a32)
    out[one_ulp] = b32[one_ulp]
    return out


# Per-scene weight tolerance.  The reference evaluates a Gaussian TransmissionSpectrum on its raw float32
# wavelengths (spectrum.py:113 + NEP 50), i.e. with numpy's float32 exp; the engine rounds a float64 exp to
# float32.  Both are float32-accurate, they differ by at most ~2 float32 ulp (parity notes).
W_RTOL = {"zoo_analytic": 3e-7}
