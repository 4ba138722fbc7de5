"""GPU parity of RenderImage.get (join-bins rescaling + colour conversions on the device)
against outputs of the reference itself (tests/golden/images_*.npz) and against the pinned numpy oracle.

Tolerance: 1e-9 relative to the image maximum (fp64 throughout; pow / atan2 / tan differ from numpy's libm in the
last bits).  The out-of-gamut mask must be identical."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
SCENES = ["double_gauss", "image_render", "arizona_eye", "spherical_aberration"]


@pytest.fixture(scope="module")
def ot():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import optrace_b200 as ot
    from optrace_b200 import engine
    engine.ensure_init()
    return ot


def _image(ot, scene):
    g = gu.load(scene)
    shape = tuple(int(v) for v in g["det0_shape"])
    data = np.zeros(shape)
    data[g["det0_yi"], g["det0_xi"]] = g["det0_vals"]
    img = ot.RenderImage(extent=g["det0_extent"])
    img.extent = np.array(g["det0_extent"], dtype=np.float64)
    img._data = data
    return img, data, dict(np.load(gu.GOLDEN / f"images_{scene}.npz"))


def _close(a, b):
    return np.all(np.abs(a - b) <= 1e-9*max(1e-300, np.abs(b).max()))


@pytest.mark.parametrize("scene", SCENES)
def test_get_matches_reference(ot, scene):
    img, data, gi = _image(ot, scene)
    N = int(gi["N"])
    for k, mode in enumerate(gi["modes"]):
        res = img.get(str(mode), N)
        r = gi[f"m{k}"]
        assert isinstance(res, ot.RGBImage if r.ndim == 3 else ot.ScalarImage)
        assert res.quantity == str(mode) and np.array_equal(res.extent, img.extent)
        d = res.data
        assert d.shape == r.shape
        if str(mode) == "Outside sRGB Gamut":
            assert np.array_equal(d, r)
        else:
            assert _close(d, r), (mode, np.abs(d - r).max())
    assert _close(img.get("sRGB (Perceptual RI)", N, L_th=0.01).data, gi["perc_lth"])
    assert _close(img.get("sRGB (Perceptual RI)", N, chroma_scale=0.5).data, gi["perc_cs"])
    assert _close(img.get("sRGB (Absolute RI)", 189).data, gi["abs_full"])


def test_get_full_resolution_and_errors(ot):
    """factor 1 (N = 945) against the oracle; argument checks of render_image.py:156-160, 219-220"""
    from oracle import image_oracle as io
    img, data, gi = _image(ot, "image_render")
    for mode in ("Irradiance", "sRGB (Absolute RI)", "Lightness (CIELUV)"):
        assert _close(img.get(mode, 945).data, io.get(data, img.extent, mode, 945))
    with pytest.raises(ValueError):
        img.get("Irradiance", 0)
    with pytest.raises(ValueError):
        img.get("no such mode")
    with pytest.raises(RuntimeError):
        ot.RenderImage(extent=[0, 1, 0, 1]).get("Irradiance")


def test_get_after_trace(ot):
    """the deliverable of BASELINE configs[1]: trace -> detector image -> RGB image, all on the device"""
    import scenes
    from oracle import image_oracle as io
    RT = scenes.double_gauss(ot)
    ot.global_options.show_warnings = False
    RT.trace(200_000)
    im = RT.detector_image()
    rgb = im.get("sRGB (Absolute RI)", 189)
    assert isinstance(rgb, ot.RGBImage) and rgb.shape[2] == 3 and 0 <= rgb.data.min() and rgb.data.max() <= 1
    assert _close(rgb.data, io.get(im.data, im.extent, "sRGB (Absolute RI)", 189))
    irr = im.get("Irradiance", 945)
    assert abs(irr.data.sum()*irr.Apx - im.power()) <= 1e-9*im.power()
