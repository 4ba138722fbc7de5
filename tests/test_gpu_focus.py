"""GPU parity of Raytracer.focus_search: per-ray work on the device (section selection,
weighted moments, cost images) against values computed by the reference itself on the same stored rays
(tests/golden/focus_*.npz) and against the pinned oracle."""
import numpy as np
import pytest

import golden_util as gu
import scenes

pytestmark = pytest.mark.gpu
Z_START = {"spherical_aberration": 23.0, "double_gauss": 120.0}


@pytest.fixture(scope="module")
def ot():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import optrace_b200 as ot
    from optrace_b200 import engine
    engine.ensure_init()
    ot.global_options.show_warnings = False
    return ot


def _traced(ot, scene):
    g = gu.load(scene)
    RT = scenes.SCENES[scene](ot)
    p0, s0, pol0, w0, wl, hz = gu.bundle(g)
    RT.trace_rays(p0, s0, pol0, w0, wl, hurb_z=hz, N_list=g["N_list"])
    return RT, g, dict(np.load(gu.GOLDEN / f"focus_{scene}.npz"))


@pytest.mark.parametrize("scene", ["spherical_aberration", "double_gauss"])
def test_cost_functions_match_reference(ot, scene):
    from optrace_b200 import engine
    RT, g, gf = _traced(ot, scene)
    L = engine.FocusLines(RT._scene.lib, RT.rays._dev, 0, RT.rays.N, float(gf["bounds"][0]) + RT.N_EPS)
    assert L.n_use == int(gf["N_use"])
    for k, m in enumerate(gf["methods"]):
        c = np.array([RT._focus_cost(L, float(z), str(m)) for z in gf["zs"]])
        assert np.allclose(c, gf[f"cost{k}"], rtol=1e-9, atol=0), (m, c, gf[f"cost{k}"])


@pytest.mark.parametrize("scene", ["spherical_aberration", "double_gauss"])
def test_focus_search_rms_matches_reference(ot, scene):
    RT, g, gf = _traced(ot, scene)
    res, info = RT.focus_search("RMS Spot Size", Z_START[scene], return_cost=True)
    assert np.allclose(info["bounds"], gf["bounds"], rtol=0, atol=1e-12) and info["N"] == int(gf["N_use"])
    # the direct solution is a quotient of two sums with cancelling terms: the summation order (block tree on the
    # device, pairwise in numpy) shows at the 1e-10 level in z and, through the mean ray slope, in the lateral position
    assert abs(res.x - float(gf["rms_x"])) <= 1e-8*abs(res.x) and abs(res.fun - float(gf["rms_fun"])) <= 1e-9*res.fun
    assert np.allclose(info["pos"], gf["rms_pos"], rtol=1e-8, atol=1e-9)
    assert info["z"].shape == (320,) and abs(info["cost"].min() - res.fun) < 0.05*res.fun + 1e-6


def test_focus_search_image_methods_and_errors(ot):
    """the optimiser-driven methods end near the RMS focus of a simple lens; argument checks of the reference"""
    RT = scenes.spherical_aberration(ot)
    RT.trace(400_000)
    rms, _ = RT.focus_search("RMS Spot Size", 23.0)
    for m in ("Irradiance Variance", "Image Sharpness", "Image Center Sharpness"):
        res, info = RT.focus_search(m, 23.0)
        assert info["bounds"][0] <= res.x <= info["bounds"][1] and abs(res.x - rms.x) < 6.0, (m, res.x, rms.x)
    with pytest.raises(ValueError):
        RT.focus_search("no such method", 23.0)
    with pytest.raises(ValueError):
        RT.focus_search("RMS Spot Size", 1e9)
    with pytest.raises(IndexError):
        RT.focus_search("RMS Spot Size", 23.0, source_index=-1)
