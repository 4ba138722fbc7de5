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


@pytest.mark.parametrize("scene", ["spherical_aberration", "image_render"])
def test_resolution_filter_matches_reference(ot, scene):
    """RenderImage.render(p, w, wl, limit) — binning + Airy-disc convolution on the device —
    against the reference's output (FFT convolution on the host); direct vs FFT convolution differ by rounding"""
    g = gu.load(scene)
    gf = dict(np.load(gu.GOLDEN / f"filter_{scene}.npz"))
    limit = float(gf["limit"])
    img = ot.RenderImage(extent=g["det0_extent0"])
    img.render(g["det0_ph"], g["det0_w"], g["det0_wl"], limit=limit)
    assert img.limit == limit and np.allclose(img.extent, gf["extent"], rtol=0, atol=1e-12)
    d = img.data
    assert d.shape == tuple(int(v) for v in gf["shape"]) and d.min() >= 0
    y0, x0 = [int(v) for v in gf["crop_origin"]]
    scale = np.abs(gf["crop"]).max(axis=(0, 1))
    assert np.all(np.abs(d[y0:y0 + 96, x0:x0 + 96] - gf["crop"]) <= 1e-9*scale)
    assert np.allclose(d.sum(axis=(0, 1)), gf["sums"], rtol=1e-9)
    assert _close(img.get("Irradiance", 189).data, gf["irr"])
    # unfiltered render + explicit filter = filtered render; projected images refuse the filter
    img2 = ot.RenderImage(extent=g["det0_extent0"])
    img2.render(g["det0_ph"], g["det0_w"], g["det0_wl"], limit=limit, _dont_filter=True)
    assert abs(img2.power() - float(np.sum(g["det0_w"].astype(np.float64)))) < 1e-9
    img2._apply_rayleigh_filter()
    assert np.allclose(img2.data, d, rtol=1e-11, atol=1e-14*d.max())      # atomic accumulation order differs between two renders
    img3 = ot.RenderImage(extent=g["det0_extent0"], projection="Equidistant")
    with pytest.raises(RuntimeError):
        img3.render(g["det0_ph"], g["det0_w"], g["det0_wl"], limit=limit)


def test_detector_image_and_iterative_render_with_limit(ot):
    """limit= through Raytracer.detector_image / iterative_render: power is conserved by the normalised kernel
    (up to what is convolved out of the enlarged extent), the extent grows by 2.7 limit, chunks are filtered once"""
    import scenes
    RT = scenes.spherical_aberration(ot)
    ot.global_options.show_warnings = False
    RT.trace(400_000)
    plain = RT.detector_image()
    filt = RT.detector_image(limit=40.0)
    assert np.allclose(filt._extent0, plain._extent0)
    assert np.allclose(filt.extent - plain.extent, np.array([-1, 1, -1, 1])*2.7*40.0/1000, atol=1e-12)
    assert abs(filt.power() - plain.power()) < 1e-6*plain.power()
    assert filt.data[:, :, 3].max() < plain.data[:, :, 3].max()           # blurred
    RT.ITER_RAYS_STEP = 200_000
    ims = RT.iterative_render(400_000, limit=40.0)
    assert ims[0].limit == 40.0 and abs(ims[0].power() - plain.power()) < 2e-2*plain.power()
