"""GPU tests of the on-device ray generator (otb_generate_rays): statistical agreement with the distributions
RaySource.create_rays draws from (ray_source.py:204-437, random.py, light_spectrum.py:81-138, srgb.py:513-553).
The reference itself only tests these statistically (tests/test_geometry.py:475-660, tests/test_misc.py:294-360)."""
import numpy as np
import pytest

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


def _gen(ot, RS, N, no_pol=False, seed=5):
    RT = ot.Raytracer(outline=[-1e5, 1e5, -1e5, 1e5, -1e5, 1e5], no_pol=no_pol)
    RT.add(RS)
    rays = RT._generate(np.array([N]), 0, N, seed)
    f = lambda t, k: t.cpu().numpy().reshape((N, k), order="F") if k > 1 else t.cpu().numpy()
    return f(rays.p0, 3), f(rays.s0, 3), None if no_pol else f(rays.pol0, 3).astype(np.float64), f(rays.w0, 1), f(rays.wl, 1)


def test_stratified_positions_and_weights(ot):
    N = 400_000
    for surf, area_test in ((ot.CircularSurface(r=2), lambda x, y: np.hypot(x, y) <= 2 + 1e-12),
                            (ot.RingSurface(r=3, ri=1.5), lambda x, y: (np.hypot(x, y) >= 1.5 - 1e-12) & (np.hypot(x, y) <= 3 + 1e-12)),
                            (ot.RectangularSurface(dim=[4, 2]), lambda x, y: (np.abs(x) <= 2) & (np.abs(y) <= 1))):
        p, s, pol, w, wl = _gen(ot, ot.RaySource(surf, pos=[1, -2, 3], power=2.5), N)
        x, y = p[:, 0] - 1, p[:, 1] + 2
        assert np.all(area_test(x, y)) and np.all(p[:, 2] == 3)
        assert np.allclose(s, [0, 0, 1]) and w.dtype == np.float32
        assert abs(float(w.astype(np.float64).sum()) - 2.5) < 1e-3
        # uniform area density: quadrant counts agree to a few per mille thanks to stratification
        q = np.array([np.count_nonzero((x > 0) == a) for a in (True, False)] + [np.count_nonzero((y > 0) == a) for a in (True, False)])
        assert np.max(np.abs(q/N - 0.5)) < 3e-3
        # mean radius squared of a uniform disc/ring/rect
        r2 = np.mean(x**2 + y**2)
        exp = {"CircularSurface": 2.0, "RingSurface": (3**2 + 1.5**2)/2, "RectangularSurface": (4**2 + 2**2)/12}[type(surf).__name__]
        assert abs(r2/exp - 1) < 2e-3
    # Line and Point
    p, *_ = _gen(ot, ot.RaySource(ot.Line(r=2, angle=30), pos=[0, 0, 0]), 100_000)
    t = p[:, 0]*np.cos(np.pi/6) + p[:, 1]*np.sin(np.pi/6)
    assert np.allclose(-p[:, 0]*np.sin(np.pi/6) + p[:, 1]*np.cos(np.pi/6), 0, atol=1e-12)
    assert abs(np.mean(t)) < 2e-3 and abs(np.std(t) - 4/np.sqrt(12)) < 2e-3
    # exact stratification in 1-D: every one of the N cells of the uniform polarisation angle is used exactly once
    _, _, pol, _, _ = _gen(ot, ot.RaySource(ot.Point(), polarization="Uniform"), 100_000)
    ang = np.mod(np.arctan2(pol[:, 1], pol[:, 0]), 2*np.pi)
    cells = np.floor(ang/(2*np.pi)*100_000).astype(int)
    assert np.count_nonzero(np.bincount(np.clip(cells, 0, 99_999), minlength=100_000) == 1) > 99_000


def test_wavelength_sampling(ot):
    from optrace_b200 import color
    N = 1_000_000
    LS = ot.LightSpectrum
    # D65 (Function spectrum -> 10000-point inverse CDF)
    *_, wl = _gen(ot, ot.RaySource(ot.Point(), spectrum=ot.presets.light_spectrum.d65), N)
    assert wl.dtype == np.float32 and wl.min() >= 380 and wl.max() <= 780
    hist, edges = np.histogram(wl, bins=40, range=(380, 780))
    mid = 0.5*(edges[1:] + edges[:-1])
    ref = color.d65_illuminant(mid)
    assert np.max(np.abs(hist/hist.sum() - ref/ref.sum())) < 1.5e-3
    # monochromatic, rectangle, lines, gaussian, blackbody
    *_, wl = _gen(ot, ot.RaySource(ot.Point(), spectrum=LS("Monochromatic", wl=532.5)), 1000)
    assert np.all(wl == np.float32(532.5))
    *_, wl = _gen(ot, ot.RaySource(ot.Point(), spectrum=LS("Rectangle", wl0=450, wl1=650)), N)
    assert wl.min() >= 450 and wl.max() <= 650 and abs(wl.mean() - 550) < 0.1 and abs(wl.std() - 200/np.sqrt(12)) < 0.1
    *_, wl = _gen(ot, ot.RaySource(ot.Point(), spectrum=LS("Lines", lines=[450, 550, 650], line_vals=[1, 2, 1])), N)
    u, c = np.unique(wl, return_counts=True)
    assert list(u) == [450, 550, 650] and np.max(np.abs(c/N - [0.25, 0.5, 0.25])) < 1e-3
    *_, wl = _gen(ot, ot.RaySource(ot.Point(), spectrum=LS("Gaussian", mu=550, sig=30)), N)
    assert abs(wl.mean() - 550) < 0.1 and abs(wl.std() - 30) < 0.1
    *_, wl = _gen(ot, ot.RaySource(ot.Point(), spectrum=LS("Blackbody", T=5500)), N)
    hist, _ = np.histogram(wl, bins=40, range=(380, 780))
    ref = color.normalized_blackbody(mid, 5500)
    assert np.max(np.abs(hist/hist.sum() - ref/ref.sum())) < 1.5e-3


def test_divergence_orientation_polarisation(ot):
    N = 500_000
    # isotropic cone: pdf(theta) ~ sin(theta) up to the cone angle -> cos(theta) uniform
    a = 20.0
    p, s, pol, w, wl = _gen(ot, ot.RaySource(ot.Point(), divergence="Isotropic", div_angle=a), N)
    # ray_source.py:312-316: theta = arccos(1 - r^2) with r^2 uniform in [0, sin^2(div_angle)] -> cos(theta) uniform
    ct = s[:, 2]
    assert ct.min() >= 1 - np.sin(np.radians(a))**2 - 1e-12
    assert abs(ct.mean() - (1 - np.sin(np.radians(a))**2/2)) < 1e-4
    assert np.allclose(np.linalg.norm(s, axis=1), 1, atol=1e-12)
    assert np.max(np.abs(np.sum(pol*s, axis=1))) < 1e-6            # pol perpendicular to s (tests/test_tracer.py:1193)
    phi = np.arctan2(s[:, 1], s[:, 0])
    assert abs(np.mean(np.cos(phi))) < 3e-3 and abs(np.mean(np.sin(phi))) < 3e-3
    # Lambertian: pdf ~ sin*cos -> sin^2(theta) uniform in [0, sin^2(a)]
    _, s, *_ = _gen(ot, ot.RaySource(ot.Point(), divergence="Lambertian", div_angle=a), N)
    st2 = 1 - s[:, 2]**2
    assert abs(st2.mean() - np.sin(np.radians(a))**2/2) < 1e-4
    # 2-D isotropic divergence in the plane of the axis angle
    _, s, *_ = _gen(ot, ot.RaySource(ot.Point(), divergence="Isotropic", div_angle=a, div_2d=True, div_axis_angle=90), N)
    assert np.max(np.abs(s[:, 0])) < 1e-12 and abs(np.mean(s[:, 1])) < 2e-3
    th = np.arccos(np.clip(s[:, 2], -1, 1))
    assert abs(th.mean() - np.radians(a)/2) < 1e-4
    # converging orientation + constant polarisation
    RS = ot.RaySource(ot.CircularSurface(r=3), orientation="Converging", conv_pos=[0, 0, 50], polarization="y")
    p, s, pol, *_ = _gen(ot, RS, 100_000)
    d = np.array([0, 0, 50.0]) - p
    assert np.allclose(s, d/np.linalg.norm(d, axis=1)[:, None], atol=1e-14)
    assert np.max(np.abs(np.sum(pol*s, axis=1))) < 1e-6 and np.all(np.abs(pol[:, 1]) > 0.99)
    # list polarisation, negative-direction error
    _, s, pol, *_ = _gen(ot, ot.RaySource(ot.Point(), polarization="List", pol_angles=[0, 90], pol_probs=[1, 3]), 100_000)
    assert abs(np.mean(np.abs(pol[:, 1]) > 0.5) - 0.75) < 2e-3
    RT = ot.Raytracer(outline=[-10, 10, -10, 10, -10, 10])
    RT.add(ot.RaySource(ot.Point(), divergence="Isotropic", div_angle=80, s=[1, 0, 0.2]))
    with pytest.raises(RuntimeError):
        RT.trace(10000)


def test_image_source(ot):
    """pixel choice follows linear pixel power; colour channel choice follows the linear RGB mixing ratios"""
    from optrace_b200 import color
    rng = np.random.default_rng(3)
    img = rng.random((12, 16, 3))
    img[3:5, 4:9] = 0            # black pixels never emit
    RS = ot.RaySource(ot.RGBImage(img, [4, 3]), pos=[0, 0, 0])
    N = 2_000_000
    p, s, pol, w, wl = _gen(ot, RS, N)
    assert np.abs(p[:, 0]).max() <= 2 and np.abs(p[:, 1]).max() <= 1.5
    PX = np.clip(np.floor((p[:, 0] + 2)/4*16).astype(int), 0, 15)
    PY = np.clip(np.floor((p[:, 1] + 1.5)/3*12).astype(int), 0, 11)
    cnt = np.zeros((12, 16))
    np.add.at(cnt, (PY, PX), 1)
    pw = color.power_from_srgb_linear(color.srgb_to_srgb_linear(img))
    assert np.all(cnt[3:5, 4:9] == 0)
    assert np.max(np.abs(cnt/N - pw/pw.sum())) < 4e-4
    # wavelengths of a pure red / green / blue image fall under the respective primary
    for c, (lo, hi) in enumerate(((560, 780), (480, 610), (380, 560))):
        mono = np.zeros((4, 4, 3))
        mono[:, :, c] = 1.0
        *_, wl = _gen(ot, ot.RaySource(ot.RGBImage(mono, [1, 1])), 200_000)
        assert np.mean((wl >= lo) & (wl <= hi)) > 0.93, c
    g = ot.RaySource(ot.GrayscaleImage(rng.random((8, 8)), [2, 2]), spectrum=ot.LightSpectrum("Monochromatic", wl=600.))
    p, _, _, _, wl = _gen(ot, g, 50_000)
    assert np.all(wl == 600) and np.abs(p[:, :2]).max() <= 1


def test_orientation_function(ot):
    """orientation="Function" (ray_source.py:274-276): or_func(x, y) translated to a device function; the directions
    the generator produces are exactly or_func evaluated at the generated positions"""
    import scenes
    RT = scenes.or_source(ot)           # engine variant with the callable: prebuilt by __graft_entry__.build()
    N = 100_000
    rays = RT._generate(np.array([N]), 0, N, 5)
    p = rays.p0.cpu().numpy().reshape((N, 3), order="F")
    s = rays.s0.cpu().numpy().reshape((N, 3), order="F")
    pol = rays.pol0.cpu().numpy().reshape((N, 3), order="F").astype(np.float64)
    ref = scenes.or_func_cone(p[:, 0], p[:, 1], f=25.0)
    assert np.max(np.abs(s - ref)) < 1e-15
    assert np.max(np.abs(np.sum(pol*s, axis=1))) < 1e-6
    # and through a trace: every ray ends on the outline end plane where its straight line says
    RT.trace(50_000)
    P = RT.rays.p_list
    d = P[:, 1] - P[:, 0]
    d /= np.linalg.norm(d, axis=1)[:, None]
    assert np.max(np.abs(d - scenes.or_func_cone(P[:, 0, 0], P[:, 0, 1], f=25.0))) < 1e-12
    assert RT.detector_image().power() > 0.99
    with pytest.raises(NotImplementedError):
        bad = ot.RaySource(ot.Point(), orientation="Function", or_func=lambda x, y: np.random.rand(x.shape[0], 3))
        RT2 = ot.Raytracer(outline=[-10, 10, -10, 10, -1, 30])
        RT2.add(bad)
        RT2.trace(1000)


@pytest.mark.parametrize("name", ["double_gauss", "arizona_eye", "image_render", "hurb_square"])
def test_fused_generation_is_bit_identical(ot, name):
    """OtbRays.gen_h: rays drawn inside the trace kernel are the rays the generator kernel writes (same device
    function, same Philox counters), so the whole ray storage and the messages are bit-identical"""
    import scenes
    res = []
    for fused in (False, True):
        RT = scenes.SCENES[name](ot)
        RT.fused_generation = fused
        RT.trace(300_005)        # divisible by the 5 sources of double_gauss: no random remainder split
        R = RT.rays
        res.append((R.p_list.copy(), R.s0_list.copy(), R.w_list.copy(), R.n_list.copy(), R.wl_list.copy(),
                    None if RT.no_pol else R.pol_list.copy(), RT._msgs.copy()))
    for a, b in zip(*res):
        if a is not None:
            assert np.array_equal(a, b, equal_nan=True)
    # the fused render path (iterative_render) as well: same images from both generation modes
    imgs = []
    for fused in (False, True):
        RT = scenes.SCENES[name](ot)
        RT.fused_generation = fused
        RT.ITER_RAYS_STEP = 100_000
        im = RT.iterative_render(200_000)[0]
        imgs.append((im.counts.copy(), im.data.copy()))
    assert np.array_equal(imgs[0][0], imgs[1][0]) and np.allclose(imgs[0][1], imgs[1][1], rtol=1e-12, atol=0)


def test_coherent_bundles_are_a_reordering(ot):
    """Raytracer.coherent_bundles hands the strata of one random variable out in blocks of 32 neighbouring cells:
    the SET of cells is the same as with the full shuffle (every cell exactly once), neighbouring rays are close in
    that variable, and the other variables stay decorrelated from it"""
    N = 320_000
    out = {}
    for coh in (False, True):
        RT = ot.Raytracer(outline=[-1e5, 1e5, -1e5, 1e5, -1e5, 1e5])
        RT.coherent_bundles = coh
        RT.add(ot.RaySource(ot.CircularSurface(r=2), divergence="Isotropic", div_angle=10, pos=[0, 0, 0]))
        rays = RT._generate(np.array([N]), 0, N, 9)
        out[coh] = (rays.p0.cpu().numpy().reshape((N, 3), order="F"), rays.s0.cpu().numpy().reshape((N, 3), order="F"))
    N2 = int(np.sqrt(N))

    def cells(s):       # invert the isotropic cone + Shirley map far enough to identify the stratum of the direction
        r = np.sqrt(np.clip(1 - s[:, 2], 0, None))/np.sin(np.radians(10))          # r in [0, 1]
        return r

    for coh in (False, True):
        p, s = out[coh]
        r = cells(s)
        # same marginal distributions: r^2 uniform (isotropic cone), uniform disc positions
        assert abs(np.mean(r**2) - 0.5) < 2e-3 and abs(np.mean(p[:, 0]**2 + p[:, 1]**2) - 2.0) < 5e-3
        # position and direction stay uncorrelated
        assert abs(np.corrcoef(p[:, 0], s[:, 0])[0, 1]) < 5e-3 and abs(np.corrcoef(p[:, 1], s[:, 1])[0, 1]) < 5e-3
    # warp coherence: the spread of the direction inside a group of 32 consecutive rays is a small fraction of the cone
    sc = out[True][1][:N - N % 32].reshape(-1, 32, 3)
    ss = out[False][1][:N - N % 32].reshape(-1, 32, 3)
    spread_c = np.median(np.ptp(sc[:, :, 0], axis=1) + np.ptp(sc[:, :, 1], axis=1))
    spread_s = np.median(np.ptp(ss[:, :, 0], axis=1) + np.ptp(ss[:, :, 1], axis=1))
    assert spread_c < 0.15*spread_s, (spread_c, spread_s)
    # the positions (not the coherent variable here) are not clustered
    pc = out[True][0][:N - N % 32].reshape(-1, 32, 3)
    assert np.median(np.ptp(pc[:, :, 0], axis=1)) > 2.0


def test_spectra_on_device_match_host_render(ot):
    """Raytracer.detector_spectrum / source_spectrum (raytracer.py:1100-1132, 1311-1329) binned on the device
    against LightSpectrum.render (light_spectrum.py:40-79 restated on the host) on the downloaded hits"""
    import scenes
    from oracle import spectrum_oracle as so
    RT = scenes.spherical_aberration(ot)
    RT.trace(400_000)
    for src in (None, 1):
        spec = RT.detector_spectrum(0, source_index=src)
        hx, hy, hw, wl, *_ = RT._hit_detector(0, src)
        m = (hw > 0).cpu().numpy()
        vals, wls = so.render(wl.cpu().numpy()[m], hw.cpu().numpy()[m])
        assert spec._wls.dtype == wls.dtype and np.array_equal(spec._wls, wls)
        assert spec._vals.shape == vals.shape and np.allclose(spec._vals, vals, rtol=2e-6, atol=1e-12*vals.max())
        assert abs(float(np.sum(spec._vals.astype(np.float64))*(wls[1] - wls[0])) - float(hw.sum())) < 1e-5*float(hw.sum())
    spec = RT.source_spectrum(1)
    b, e = RT.rays._local_range(1)
    vals, wls = so.render(RT.rays.wl_list[b:e], RT.rays.w_list[b:e, 0])
    assert np.array_equal(spec._wls, wls) and np.allclose(spec._vals, vals, rtol=2e-6, atol=1e-12*vals.max())
    # monochromatic source: the +-1 nm window of light_spectrum.py:66-68
    RT2 = ot.Raytracer(outline=[-5, 5, -5, 5, -1, 10])
    RT2.add(ot.RaySource(ot.CircularSurface(r=1), spectrum=ot.LightSpectrum("Monochromatic", wl=555.0), pos=[0, 0, 0]))
    RT2.add(ot.Detector(ot.RectangularSurface(dim=[4, 4]), pos=[0, 0, 5]))
    RT2.trace(10_000)
    spec = RT2.detector_spectrum()
    vals, wls = so.render(RT2.rays.wl_list, RT2.rays.w_list[:, 0])
    assert np.array_equal(spec._wls, wls) and np.allclose(spec._vals, vals, rtol=2e-6)


@pytest.mark.parametrize("preset", ["tv_testcard2", "ETDRS_chart_inverted"])
def test_reference_image_presets_pixel_sampling(ot, preset):
    """C3 / C4 sources (examples/arizona_eye_model.py, image_render_many_rays.py) on the reference's OWN images
    (optrace_b200/data/images = optrace/resources/images): generated positions follow the linear-light pixel power
    (ray_source.py:237-255, random.py:129-140), pixel by pixel; RGB images also the channel mixing (srgb.py:513-553)"""
    from optrace_b200 import color
    im = getattr(ot.presets.image, preset)([4, 3])
    RS = ot.RaySource(im, pos=[0, 0, 0])
    N = 4_000_000
    p, s, pol, w, wl = _gen(ot, RS, N)
    H, W = im.shape[:2]
    PX = np.clip(np.floor((p[:, 0] + 2)/4*W).astype(int), 0, W - 1)
    PY = np.clip(np.floor((p[:, 1] + 1.5)/3*H).astype(int), 0, H - 1)
    cnt = np.zeros((H, W))
    np.add.at(cnt, (PY, PX), 1)
    data = im.data
    if data.ndim == 3:
        pw = color.power_from_srgb_linear(color.srgb_to_srgb_linear(data))
    else:
        pw = color.srgb_to_srgb_linear(data)          # grayscale: pixel value through the sRGB curve (ray_source.py:120-144)
    pw = pw/pw.sum()
    assert np.all(cnt[pw == 0] == 0)                   # black pixels never emit
    # stratified inverse-CDF sampling: a pixel owns an interval of the CDF axis, the N strata cut it with at most one
    # partial stratum at either end — every count within two of its expectation N p (plain sampling: ~sqrt(N p))
    assert np.max(np.abs(cnt - N*pw)) <= 2.0 + 1e-6
    assert np.mean(np.abs(cnt - N*pw)[pw > 0]) < 0.6
    if data.ndim == 3:
        # wavelengths: mean power share of the three primaries over the whole image
        lin = color.srgb_to_srgb_linear(data)
        share = np.array([color.SRGB_R_PRIMARY_POWER_FACTOR*lin[..., 0].sum(), color.SRGB_G_PRIMARY_POWER_FACTOR*lin[..., 1].sum(),
                          color.SRGB_B_PRIMARY_POWER_FACTOR*lin[..., 2].sum()])
        share /= share.sum()
        # the primaries overlap: compare the mean wavelength with the mixture of the primaries' mean wavelengths
        wl5, Fr, Fg, Fb = color.srgb_primary_cdfs()
        means = [np.sum(0.5*(wl5[1:] + wl5[:-1])*np.diff(F))/F[-1] for F in (Fr, Fg, Fb)]
        assert abs(wl.mean() - float(np.dot(share, means))) < 0.3
