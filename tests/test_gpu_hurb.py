"""HURB acceptance statistics on the DEVICE-RNG path (Philox + Box-Muller deviates drawn inside the trace kernel,
otb_trace.cu store_step): the reference's own assertions (tests/test_tracer_hurb.py:19-141) with the scenes of its
tests/hurb_geometry.py, restated through the drop-in API.  The golden fixtures cover HURB with injected deviates bit
for bit; these tests cover what the fixtures cannot: that the on-device normal deviates give the diffraction profiles
the reference accepts (sigma ratios against the analytic Airy / sinc^2 / Fresnel-edge curves)."""
import numpy as np
import pytest
import scipy.special
import scipy.ndimage
from scipy.interpolate import RectBivariateSpline

pytestmark = pytest.mark.gpu


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


def _airy(x, wl, n, r, z):
    Rnz = 2*np.pi/(wl*1e-9)*n*r/z*x*1e-3
    return (2*scipy.special.j1(Rnz)/Rnz)**2


def _radial_profile(img, N_px):
    imgi = img.get("Irradiance", N_px)
    bins, c1 = imgi.profile(x=0)
    bins, c2 = imgi.profile(y=0)
    c = 0.5*(c1[0] + c2[0])
    return bins[:-1] + (bins[1] - bins[0])/2, c/np.max(c)


def hurb_pinhole(ot, n, ri, wl, zd, N, N_px, dim_ext_fact, lens=False, use_hurb=True):
    """tests/hurb_geometry.py:87-154 (pinhole) and :10-84 (aperture directly before an ideal lens)"""
    RT = ot.Raytracer(outline=[-15, 15, -15, 15, -6, zd + 10], use_hurb=use_hurb, n0=ot.RefractionIndex("Constant", n))
    RT.HURB_FACTOR = 1
    RT.add(ot.RaySource(ot.CircularSurface(r=ri), s=[0, 0, 1], pos=[0, 0, -5], spectrum=ot.LightSpectrum("Monochromatic", wl=wl)))
    if lens:
        RT.add(ot.Aperture(ot.RingSurface(r=ri + 1, ri=ri), pos=[0, 0, -0.001]))
        RT.add(ot.IdealLens(ri + 1, 1/zd*1000, pos=[0, 0, 0]))
    else:
        RT.add(ot.Aperture(ot.RingSurface(r=ri + 5, ri=ri), pos=[0, 0, 0]))
    dim_ext = 1.22/(2*np.pi/(wl*1e-9)*n*ri/zd/np.pi)*1e3*6*dim_ext_fact
    RT.add(ot.Detector(ot.RectangularSurface(dim=[dim_ext, dim_ext]), pos=[0, 0, zd]))
    RT.trace(N)
    r, imgic = _radial_profile(RT.detector_image(), N_px)
    return r, imgic, _airy(r, wl, n, ri, zd)


@pytest.mark.parametrize("n, ri, wl, zd", [[1, 0.02, 550, 20], [1.33, 0.012, 380, 30], [1.5, 0.005, 780, 23], [1.1, 0.01, 480, 20]])
def test_hurb_error_pinhole(ot, n, ri, wl, zd):
    """tests/test_tracer_hurb.py:54-67"""
    r, imgi, imgr = hurb_pinhole(ot, n, ri, wl, zd, N=2_000_000, N_px=315, dim_ext_fact=3)
    std_i = np.average(r**2, weights=imgi)**0.5
    std_r = np.average(r**2, weights=imgr)**0.5
    assert abs(std_i/std_r - 0.95) < 0.04, std_i/std_r


@pytest.mark.parametrize("n, ri, wl, zd", [[1, 1, 550, 20], [1.33, 3, 380, 30], [1.5, 5, 780, 23], [1.1, 2.7, 480, 20]])
def test_hurb_error_lens(ot, n, ri, wl, zd):
    """tests/test_tracer_hurb.py:116-130"""
    r, imgi, imgr = hurb_pinhole(ot, n, ri, wl, zd, N=1_000_000, N_px=315, dim_ext_fact=3, lens=True)
    std_i = np.average(r**2, weights=imgi)**0.5
    std_r = np.average(r**2, weights=imgr)**0.5
    assert abs(std_i/std_r - 0.95) < 0.04, std_i/std_r


def test_hurb_setting(ot):
    """tests/test_tracer_hurb.py:132-141: without HURB the ideal lens focuses to a point"""
    r, imgi, imgr = hurb_pinhole(ot, 1.1, 2, 550, 20, N=100_000, N_px=945, dim_ext_fact=3, lens=True, use_hurb=False)
    assert np.average(r**2, weights=imgi)**0.5 < 1e-10


@pytest.mark.parametrize("n, d1, d2, wl, zd, ang", [[1, 0.02, 0.1, 550, 20, 0], [1.33, 0.012, 0.05, 380, 30, 10],
                                                   [1.5, 0.005, 0.005, 780, 23., -30], [1.1, 0.01, 0.1, 480, 20, 45]])
def test_hurb_error_slit(ot, n, d1, d2, wl, zd, ang):
    """tests/test_tracer_hurb.py:95-113 with the scene of tests/hurb_geometry.py:157-251 (rotated slit)"""
    N, N_px, dim_ext_fact = 5_000_000, 945, 5
    slit = lambda x, d: np.sinc(d*1e-3*n/(wl*1e-9)*x/zd)**2            # noqa: E731
    dim_ext = 5/(min(d1, d2)*1e-3*n/(wl*1e-9)/zd)*dim_ext_fact
    RT = ot.Raytracer(outline=[-dim_ext, dim_ext, -dim_ext, dim_ext, -6, zd + 10], use_hurb=True, n0=ot.RefractionIndex("Constant", n))
    RT.HURB_FACTOR = 1
    RS = ot.RaySource(ot.RectangularSurface(dim=[d1, d2]), s=[0, 0, 1], pos=[0, 0, -5], spectrum=ot.LightSpectrum("Monochromatic", wl=wl))
    RS.rotate(ang)
    RT.add(RS)
    ap = ot.Aperture(ot.SlitSurface(dim=[d1 + 2, d2 + 2], dimi=[d1, d2]), pos=[0, 0, 0])
    ap.rotate(ang)
    RT.add(ap)
    RT.add(ot.Detector(ot.RectangularSurface(dim=[dim_ext, dim_ext]), pos=[0, 0, zd]))
    RT.trace(N)
    img = RT.detector_image()
    imgi = img.get("Irradiance", N_px)
    x_i = np.linspace(imgi.extent[0], imgi.extent[1], N_px)
    y_i = np.linspace(imgi.extent[2], imgi.extent[3], N_px)
    interp = RectBivariateSpline(y_i, x_i, imgi.data, kx=3, ky=3)
    r = np.linspace(img.extent[0], img.extent[1], N_px)
    a = np.deg2rad(ang)
    c1 = interp(r*np.sin(a), r*np.cos(a), grid=False)
    c2 = interp(r*np.sin(a + np.pi/2), r*np.cos(a + np.pi/2), grid=False)
    c1, c2 = c1/np.max(c1), c2/np.max(c2)
    for c, d, delta in ((c1, d1, 0.05), (c2, d2, 0.09)):
        std_i = np.average(r**2, weights=c)**0.5
        std_r = np.average(r**2, weights=slit(r, d))**0.5
        assert abs(std_i/std_r - 1.11) < delta, (std_i/std_r, d)


@pytest.mark.parametrize("n, wl, zd", [[1, 550, 20], [1.33, 380, 30], [1.5, 780, 23], [1.1, 480, 20]])
def test_hurb_error_edge(ot, n, wl, zd):
    """tests/test_tracer_hurb.py:70-93 with the scene of tests/hurb_geometry.py:254-325 (straight edge)"""
    N, N_px, dim_ext_fact = 3_000_000, 945, 2.5

    def edge_curve(x):
        u_ = np.sqrt(2*n/(wl*1e-9)/(zd*1e-3))*x*1e-3
        S, C = scipy.special.fresnel(u_)
        return 0.5*((S + 0.5)**2 + (C + 0.5)**2)

    dim_ext = 0.5*2*dim_ext_fact
    RT = ot.Raytracer(outline=[-4*dim_ext, 4*dim_ext, -4*dim_ext, 4*dim_ext, -6, zd + 10], use_hurb=True,
                      n0=ot.RefractionIndex("Constant", n))
    RT.HURB_FACTOR = 1
    RT.add(ot.RaySource(ot.RectangularSurface(dim=[dim_ext/2, dim_ext/2]), s=[0, 0, 1], pos=[0, dim_ext/4, -1],
                        spectrum=ot.LightSpectrum("Monochromatic", wl=wl)))
    RT.add(ot.Aperture(ot.SlitSurface(dim=[4*dim_ext, 4*dim_ext], dimi=[4*dim_ext - 0.4, 4*dim_ext - 0.4]),
                       pos=[0, (4*dim_ext - 0.4)/2, 0]))
    RT.add(ot.Detector(ot.RectangularSurface(dim=[dim_ext, dim_ext]), pos=[0, 0, zd]))
    RT.trace(N)
    imgi = RT.detector_image().get("Irradiance", N_px)
    imgic = np.mean(imgi.data, axis=1)
    imgic /= np.mean(imgic[4*(imgic.shape[0]//5):])
    r = np.linspace(imgi.extent[2], imgi.extent[3], imgi.shape[0])
    imgr = edge_curve(r)
    ind = np.argmax(imgr > 1.2)
    imgrf = scipy.ndimage.gaussian_filter1d(imgr, sigma=10)
    assert np.sqrt(np.mean((imgrf[ind:-2] - imgic[ind:-2])**2)) < 0.02
    assert np.sqrt(np.mean((imgr[:ind]**0.5 - imgic[:ind]**0.5)**2)) < 0.015


def test_hurb_masking(ot):
    """tests/test_tracer_hurb.py:19-51: filter and aperture both remove rays before the bending"""
    RT = ot.Raytracer(outline=[-15, 15, -15, 15, -6, 10], use_hurb=True)
    RT.add(ot.RaySource(ot.CircularSurface(r=3), s=[0, 0, 1], pos=[0, 0, -5], spectrum=ot.presets.light_spectrum.d65))
    RT.add(ot.Filter(ot.CircularSurface(r=5), pos=[0, 0, -1],
                     spectrum=ot.TransmissionSpectrum("Rectangle", wl0=500, wl1=650, val=1)))
    RT.add(ot.Aperture(ot.RingSurface(r=5, ri=2.9), pos=[0, 0, 0]))
    RT.trace(200_000)
    W = RT.rays.w_list
    assert np.count_nonzero(W[:, 0]) > np.count_nonzero(W[:, 1]) > np.count_nonzero(W[:, 2])
    # bent rays keep unit directions and polarisation perpendicular to them
    s = RT.rays.s0_list
    assert np.allclose(np.linalg.norm(s, axis=1), 1, atol=1e-12)
