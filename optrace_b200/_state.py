"""Cheap structural state of scene objects (the role of BaseClass.crepr() in the reference, base_class.py:28-59):
a nested tuple of plain attribute values that changes whenever an object is moved, flipped, rotated or
re-parameterised.  Used by the Raytracer to detect geometry changes without re-flattening the scene."""
import numpy as np


def state_of(obj):
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    if isinstance(obj, np.ndarray):
        return obj.tobytes() if obj.size <= 64 else (id(obj), obj.shape)
    if isinstance(obj, (list, tuple)):
        return tuple(state_of(v) for v in obj)
    if isinstance(obj, dict):
        return tuple((k, state_of(v)) for k, v in obj.items())
    if isinstance(obj, (np.floating, np.integer)):
        return obj.item()
    if hasattr(obj, "__dict__") and type(obj).__module__.startswith("optrace_b200"):
        return (type(obj).__name__,) + tuple((k, state_of(v)) for k, v in obj.__dict__.items() if k not in ("desc", "long_desc"))
    return id(obj)      # callables, scipy spline objects, images


# Scene epoch: bumped by every attribute assignment on a scene object (surfaces, elements, media, spectra).
# Lets the Raytracer reuse its last structural state when nothing was assigned since (the deep walk above costs
# ~0.1 ms and sits on the critical path between a synchronisation point and the next kernel launch).
EPOCH = [0]


def touch(*_):
    EPOCH[0] += 1
