"""ctypes mirror of include/otb.h and loader of the CUDA engine (libotb.so).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no CPU fallback:
if the shared object is missing, `lib()` raises with the build instruction.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib

NPAR = 20
NMSG = 5

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("OTB_LIB", _HERE / "csrc" / "libotb.so"))   # OTB_LIB: experiment builds


class OtbSurface(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("func_id", C.c_int32), ("aux_off", C.c_int32),
                ("aux_n0", C.c_int32), ("aux_n1", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32),
                ("pos", C.c_double*3), ("r", C.c_double), ("z_min", C.c_double), ("z_max", C.c_double),
                ("par", C.c_double*NPAR)]


class OtbMedium(C.Structure):
    _fields_ = [("model", C.c_int32), ("func_id", C.c_int32), ("aux_off", C.c_int32), ("aux_n", C.c_int32),
                ("c", C.c_double*12)]


class OtbFilter(C.Structure):
    _fields_ = [("type", C.c_int32), ("inverse", C.c_int32), ("func_id", C.c_int32), ("aux_off", C.c_int32),
                ("aux_n", C.c_int32), ("pad", C.c_int32), ("c", C.c_double*4)]


class OtbStep(C.Structure):
    _fields_ = [("role", C.c_int32), ("surface", C.c_int32), ("medium_after", C.c_int32), ("filter", C.c_int32),
                ("hurb", C.c_int32), ("hurb_slot", C.c_int32), ("D", C.c_double)]


class OtbSceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_surfaces", C.c_int32), ("n_steps", C.c_int32),
                ("n_media", C.c_int32), ("n_filters", C.c_int32), ("no_pol", C.c_int32),
                ("medium0", C.c_int32), ("n_hurb", C.c_int32), ("n_aux", C.c_int64),
                ("outline", C.c_double*6), ("hurb_factor", C.c_double),
                ("surfaces", C.POINTER(OtbSurface)), ("steps", C.POINTER(OtbStep)),
                ("media", C.POINTER(OtbMedium)), ("filters", C.POINTER(OtbFilter)),
                ("aux", C.POINTER(C.c_double)), ("arithmetic", C.c_int32), ("pad", C.c_int32)]


class OtbRays(C.Structure):
    _fields_ = [("N", C.c_int64), ("p0_d", C.c_void_p), ("s0_d", C.c_void_p), ("pol0_d", C.c_void_p),
                ("w0_d", C.c_void_p), ("wl_d", C.c_void_p), ("hurb_z_d", C.c_void_p),
                ("seed", C.c_uint64), ("ray_offset", C.c_int64), ("gen_h", C.c_void_p)]


class OtbRayStore(C.Structure):
    _fields_ = [("N", C.c_int64), ("nt", C.c_int32), ("pad", C.c_int32), ("p_d", C.c_void_p),
                ("s_d", C.c_void_p), ("pol_d", C.c_void_p), ("w_d", C.c_void_p), ("n_d", C.c_void_p),
                ("wl_d", C.c_void_p), ("trace_status_d", C.c_void_p)]


class OtbDetector(C.Structure):
    _fields_ = [("surface", OtbSurface), ("projection", C.c_int32), ("has_extent", C.c_int32),
                ("extent", C.c_double*4)]


class OtbSource(C.Structure):
    _fields_ = [("shape", C.c_int32), ("orientation", C.c_int32), ("divergence", C.c_int32),
                ("polarization", C.c_int32), ("wl_mode", C.c_int32), ("div_2d", C.c_int32),
                ("img_w", C.c_int32), ("img_h", C.c_int32),
                ("n_rays", C.c_int64), ("ray_start", C.c_int64), ("gid_start", C.c_int64),
                ("power", C.c_double), ("weight", C.c_double), ("pos", C.c_double*3),
                ("geom", C.c_double*8), ("extent", C.c_double*4), ("s", C.c_double*3),
                ("conv_pos", C.c_double*3), ("div_sin", C.c_double), ("div_angle", C.c_double),
                ("div_axis", C.c_double), ("pol_angle", C.c_double), ("wl", C.c_double*4),
                ("wl_tab_off", C.c_int32), ("wl_tab_n", C.c_int32), ("div_tab_off", C.c_int32),
                ("div_tab_n", C.c_int32), ("pol_tab_off", C.c_int32), ("pol_tab_n", C.c_int32),
                ("pix_cdf_off", C.c_int32), ("pix_cdf_n", C.c_int32), ("pix_rgb_off", C.c_int32),
                ("srgb_off", C.c_int32), ("or_func_id", C.c_int32), ("coherent", C.c_int32)]


class OtbGenerator(C.Structure):
    _fields_ = [("sources_h", C.POINTER(OtbSource)), ("n_sources", C.c_int32), ("pad", C.c_int32),
                ("gen_aux_d", C.c_void_p)]


class OtbDeviceInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_major", C.c_int32), ("sm_minor", C.c_int32),
                ("sm_count", C.c_int32), ("total_mem", C.c_int64), ("l2_bytes", C.c_int32),
                ("max_smem_per_block", C.c_int32), ("name", C.c_char*64)]


# every symbol include/otb.h declares: (name, restype, argtypes)
_VP, _I32, _I64, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
SYMBOLS = {
    "otb_init": (C.c_int, [C.c_int]),
    "otb_last_error": (C.c_char_p, []),
    "otb_abi_version": (C.c_int, []),
    "otb_device_info": (C.c_int, [C.POINTER(OtbDeviceInfo)]),
    "otb_dev_alloc": (C.c_int, [C.POINTER(_VP), C.c_size_t]),
    "otb_dev_free": (C.c_int, [_VP]),
    "otb_memcpy_h2d": (C.c_int, [_VP, _VP, C.c_size_t, _VP]),
    "otb_memcpy_d2h": (C.c_int, [_VP, _VP, C.c_size_t, _VP]),
    "otb_memset_d": (C.c_int, [_VP, C.c_int, C.c_size_t, _VP]),
    "otb_stream_sync": (C.c_int, [_VP]),
    "otb_scene_create": (C.c_int, [C.POINTER(OtbSceneDesc), C.POINTER(_VP)]),
    "otb_scene_destroy": (C.c_int, [_VP]),
    "otb_scene_update": (C.c_int, [_VP, C.POINTER(OtbSceneDesc), _VP]),
    "otb_focus_prepare": (C.c_int, [C.POINTER(OtbRayStore), C.c_int64, C.c_int64, C.c_double, _VP, _VP, _VP, _VP, _VP,
                                    _VP, _VP, _VP]),
    "otb_focus_moments": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_int64, C.c_int32, _VP, _VP, _VP]),
    "otb_focus_image": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_int64, C.c_double, C.c_int32, C.c_int32, _VP, _VP, _VP]),
    "otb_spectrum_stats": (C.c_int, [_VP, _VP, C.c_int64, C.c_int32, _VP, _VP, _VP]),
    "otb_spectrum_hist": (C.c_int, [_VP, _VP, C.c_int64, C.c_int32, _VP, C.c_int32, _VP, _VP]),
    "otb_image_tiles_mask": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP]),
    "otb_image_tiles_pack": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32, _VP, C.c_int32, _VP, _VP, _VP]),
    "otb_image_tiles_unpack": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32, _VP, C.c_int32, _VP, _VP]),
    "otb_image_convolve": (C.c_int, [_VP, C.c_int32, C.c_int32, _VP, C.c_int32, _VP, _VP]),
    "otb_image_rescale": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP]),
    "otb_image_stats": (C.c_int, [_VP, C.c_int64, C.c_int32, C.c_double, _VP, _VP]),
    "otb_image_convert": (C.c_int, [_VP, C.c_int64, C.c_int32, C.c_double, C.c_double, _VP, _VP, _VP]),
    "otb_trace_store": (C.c_int, [_VP, C.POINTER(OtbRays), C.POINTER(OtbRayStore), _VP, _VP, _VP]),
    "otb_generate_rays": (C.c_int, [C.POINTER(OtbSource), C.c_int, _VP, _I64, _U64, _I64, C.c_int,
                                    _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "otb_hurb_normals": (C.c_int, [_I64, _U64, _I64, _I32, _VP, _VP, _VP]),
    "otb_detector_hits": (C.c_int, [C.POINTER(OtbRayStore), _I64, _I64, C.POINTER(OtbDetector),
                                    _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "otb_render_xyzw": (C.c_int, [_VP, _VP, _VP, _VP, _I64, C.POINTER(C.c_double), _I32, _I32, _VP, _VP, _VP]),
    "otb_trace_render": (C.c_int, [_VP, C.POINTER(OtbRays), C.c_int, C.POINTER(OtbDetector),
                                   C.POINTER(C.c_double), C.POINTER(_I32), C.POINTER(_I32),
                                   C.POINTER(_VP), C.POINTER(_VP), _VP, _VP, _VP, _VP]),
    "otb_surface_find_hit": (C.c_int, [C.POINTER(OtbSurface), C.POINTER(C.c_double), _I64, _I64,
                                       _VP, _VP, _VP, _VP, _VP, _VP]),
    "otb_surface_normals": (C.c_int, [C.POINTER(OtbSurface), C.POINTER(C.c_double), _I64, _I64,
                                      _VP, _VP, _VP, _VP]),
    "otb_surface_values": (C.c_int, [C.POINTER(OtbSurface), C.POINTER(C.c_double), _I64, _I64,
                                     _VP, _VP, _VP, _VP, _VP]),
    "otb_medium_eval": (C.c_int, [C.POINTER(OtbMedium), C.POINTER(C.c_double), _I64, _I64, _VP, _VP, _VP]),
    "otb_sphere_projection": (C.c_int, [C.POINTER(OtbSurface), C.c_int, _I64, _VP, _VP, _VP]),
    "otb_selftest_division": (C.c_int, [_I64, _VP, _VP, _VP, _VP, _VP]),
}

_lib = None
_variants = {}


class EngineError(RuntimeError):
    pass


def lib(path: os.PathLike | None = None) -> C.CDLL:
    """Loads libotb.so (once) and declares the prototypes.  Fails loudly when the extension is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = pathlib.Path(path) if path is not None else LIB_PATH
    if str(p) in _variants:
        return _variants[str(p)]
    if not p.exists():
        raise EngineError(f"CUDA engine {p} is missing. Build it with `python -c 'import __graft_entry__ as g; "
                          f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    l = C.CDLL(str(p))
    for name, (res, args) in SYMBOLS.items():
        f = getattr(l, name)
        f.restype, f.argtypes = res, args
    if l.otb_abi_version() != 2:
        raise EngineError("ABI version mismatch between _cabi.py and libotb.so")
    if path is None:
        _lib = l
    _variants[str(p)] = l
    return l


def check(code: int, l: C.CDLL | None = None) -> None:
    """Maps OtbStatus to the reference's exception types (SURVEY.md §8b 'Errors')."""
    if code == 0:
        return
    l = l or lib()
    msg = l.otb_last_error().decode(errors="replace")
    if code == 1:
        raise ValueError(msg)
    if code == 6:
        raise TimeoutError(msg)
    if code == 3:
        raise MemoryError(msg)
    raise EngineError(f"[otb status {code}] {msg}")
