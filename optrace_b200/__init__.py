"""optrace_b200 — B200-native sequential raytracing engine behind the public API of drocheam/optrace.

    import optrace_b200 as ot
    RT = ot.Raytracer(outline=[...]); RT.add(ot.RaySource(...)); RT.add(ot.Lens(...)); RT.add(ot.Detector(...))
    RT.trace(10_000_000); img = RT.detector_image()

The hot path (surface intersection, Snell/Fresnel/polarisation, filters/apertures, HURB, detector binning,
ray generation) runs in hand-written sm_100a CUDA kernels reached through the C ABI of include/otb.h; this
package is the thin host side (scene model, flattening, ctypes).  There is no CPU fallback.
"""
from .options import global_options, OptraceWarning, warning  # noqa: F401
from .surfaces import (Surface, Point, Line, CircularSurface, RectangularSurface, RingSurface, SlitSurface,  # noqa: F401
                       ConicSurface, SphericalSurface, TiltedSurface, AsphericSurface, FunctionSurface1D,
                       FunctionSurface2D, DataSurface1D, DataSurface2D)
from .media import Spectrum, LightSpectrum, TransmissionSpectrum, RefractionIndex  # noqa: F401
from .images import RGBImage, GrayscaleImage, ScalarImage, RenderImage  # noqa: F401
from .elements import Element, Lens, IdealLens, Filter, Aperture, Detector, RaySource, Group, PointMarker  # noqa: F401
from .load import load_agf, load_zmx  # noqa: F401
from .ray_storage import RayStorage  # noqa: F401
from .raytracer import Raytracer  # noqa: F401
from . import presets, color  # noqa: F401

__version__ = "0.1.0"
