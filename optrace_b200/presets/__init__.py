"""Presets of the drop-in API (subset on the hot path's input side): light spectra, refraction indices,
procedural test images and the Arizona eye geometry.  Reference: optrace/tracer/presets/."""
from . import light_spectrum, refraction_index, image, geometry, spectral_lines  # noqa: F401
