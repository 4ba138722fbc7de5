"""GPU tests of the array entry points (Surface.find_hit / normals / values, RefractionIndex.__call__, sphere
projections, RenderImage.render) written like the reference's own tests (tests/test_surface.py:126-268,
tests/test_refraction_index.py, tests/test_image.py:108-173, tests/test_misc.py:139-170), plus the fused
render path against the store path."""
import numpy as np
import pytest

import golden_util as gu
import scenes

pytestmark = pytest.mark.gpu
X0, Y0, Z0 = 1.24, -5.8, 0.01


@pytest.fixture(scope="module")
def ot():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import optrace_b200 as ot
    from optrace_b200 import engine
    engine.ensure_init()
    return ot


def test_surface_normals_known_answers(ot):
    """table of tests/test_surface.py:158-174"""
    for surf, n, _ in scenes.kat_surfaces(ot):
        if n is None:
            continue
        surf.move_to([X0, Y0, Z0])
        na = surf.normals(np.array([X0 + 1]), np.array([Y0 + 0.5]))
        assert np.allclose(np.array(n) - na, 0, atol=1e-8), (type(surf).__name__, na)


def test_surface_values_known_answers(ot):
    """table of tests/test_surface.py:204-219"""
    from optrace_b200 import engine
    for surf, _, z in scenes.kat_surfaces(ot):
        if z is None:
            continue
        surf.move_to([X0, Y0, Z0])
        za, _ = engine.surface_values(surf, np.array([X0 + 1, X0]), np.array([Y0 + 0.5, Y0 - 0.5]))
        assert np.allclose(za - Z0, z, atol=1e-7), (type(surf).__name__, za - Z0)


def test_normals_properties_and_oracle(ot):
    """tests/test_surface.py:126-155 + element-wise comparison with the oracle"""
    from optrace_b200.scene import standalone_surface
    from oracle import trace_oracle as orc
    for surf in scenes.hit_test_surfaces(ot):
        x = np.linspace(surf.extent[0] - 1, surf.extent[1] + 2, 1000)
        y = np.linspace(surf.extent[2] - 1, surf.extent[3] + 2, 1000)
        n = surf.normals(x, y)
        assert np.all(n[:, 2] > 0)
        assert np.allclose(n[:, 0]**2 + n[:, 1]**2 + n[:, 2]**2 - 1, 0)
        m = surf.mask(x, y)
        assert np.all(n[~m] == [0, 0, 1])
        if surf.is_flat():
            assert np.all(n[:, 2] == 1)
        rec, aux, funcs = standalone_surface(surf)
        ref = orc.surf_normals(rec, x, y, orc.Ctx(aux, funcs))
        assert np.max(np.abs(n - ref)) < 1e-9, type(surf).__name__


def test_surface_hit_finding(ot):
    """tests/test_surface.py:237-268 + element-wise comparison with the oracle"""
    from optrace_b200 import engine
    from optrace_b200.scene import standalone_surface
    from oracle import trace_oracle as orc
    rng = np.random.default_rng(3)
    p = rng.uniform(-2, -1, size=(10000, 3))
    s = rng.uniform(-1, 1, size=(10000, 3))
    s /= np.linalg.norm(s, axis=1)[:, None]
    s[:, 2] = np.abs(s[:, 2])
    for surf in scenes.hit_test_surfaces(ot):
        name = type(surf).__name__
        p_hit, is_hit, ill = surf.find_hit(p, s)
        z_hit, _ = engine.surface_values(surf, p_hit[is_hit, 0], p_hit[is_hit, 1])
        assert np.allclose(p_hit[is_hit, 2] - z_hit, 0, rtol=0, atol=1e-6), name
        zs, ms = engine.surface_values(surf, p_hit[~is_hit, 0], p_hit[~is_hit, 1])
        assert np.all(p_hit[~is_hit, 2][ms] > zs[ms]), name
        t = (p_hit[:, 2] - p[:, 2])/s[:, 2]
        assert np.allclose(p + s*t[:, None] - p_hit, 0, atol=1e-6), name
        rec, aux, funcs = standalone_surface(surf)
        rp, rh, ri = orc.surf_find_hit(rec, p, s, orc.Ctx(aux, funcs))
        assert np.array_equal(is_hit, rh) and np.array_equal(ill, ri), name
        assert np.max(np.abs(p_hit - rp)) < 1e-9, (name, np.max(np.abs(p_hit - rp)))
        # rays starting behind the surface keep their position and never hit
        p2 = p_hit.copy()
        p2[:, 2] = surf.z_max + 2
        p_hit2, is_hit2, _ = surf.find_hit(p2, s)
        assert np.allclose(p_hit2 - p2, 0) and not np.any(is_hit2), name


def test_where_argument_and_small_inputs(ot):
    surf = ot.SphericalSurface(r=3, R=5)
    p = np.array([[0., 0., -3.], [1., 0., -3.], [9., 0., -3.]])
    s = np.array([[0., 0., 1.]]*3)
    ph, hit, ill = surf.find_hit(p, s, where=np.array([True, False, True]))
    assert ph.shape == (2, 3) and list(hit) == [True, False] and ill.shape == (2,)
    ph1, hit1, _ = surf.find_hit(p[:1], s[:1])
    assert ph1.shape == (1, 3) and hit1[0] and abs(ph1[0, 2]) < 1e-15


def test_refraction_index_models(ot):
    """known answers (BK7 n_d, Abbe numbers) and oracle agreement for every dispersion model"""
    from oracle import trace_oracle as orc
    RI = ot.RefractionIndex
    bk7 = ot.presets.refraction_index.BK7
    assert abs(float(bk7(np.array([587.5618]))[0]) - 1.5168) < 1e-4
    assert abs(bk7.abbe_number() - 64.17) < 0.05
    n = RI("Abbe", n=1.6, V=35)
    assert abs(n.abbe_number() - 35) < 1e-3 and abs(float(n(np.array([587.5618]))[0]) - 1.6) < 1e-6
    assert not RI("Constant", n=1.3).is_dispersive()
    wl = np.linspace(380, 780, 401)
    models = list(ot.presets.refraction_index.all_presets) + [
        RI("Cauchy", coeff=[1.5, 0.004, 1e-5, 1e-7]), RI("Conrady", coeff=[1.47, 0.015, 3.5e-5]),
        RI("Sellmeier2", coeff=[1.045, 0.266, 0.206, 0.001, 0.3]), RI("Sellmeier4", coeff=[1.5, 0.9, 0.01, 0.5, 100.0]),
        RI("Sellmeier5", coeff=[0.6, 0.005, 0.4, 0.01, 0.9, 100.0, 0.01, 0.02, 0.001, 0.03]),
        RI("Schott", coeff=[2.27, -0.01, 0.012, 0.0003, -1e-5, 1e-6]),
        RI("Herzberger", coeff=[1.5, 0.004, 1e-4, -0.01, 1e-3, -1e-4]),
        RI("Handbook of Optics 1", coeff=[2.2, 0.02, 0.03, 0.01]), RI("Handbook of Optics 2", coeff=[1.6, 0.9, 0.02, 0.01]),
        RI("Extended", coeff=[2.27, -0.01, 0.012, 0.0003, -1e-5, 1e-6, 1e-8, -1e-9]),
        RI("Extended2", coeff=[2.27, -0.01, 0.012, 0.0003, -1e-5, 1e-6, 1e-3, -1e-4]),
        RI("Extended3", coeff=[2.27, -0.01, 1e-3, 0.012, 0.0003, -1e-5, 1e-4, -1e-5, 1e-9])]
    for m in models:
        if m.spectrum_type == "Function":
            continue
        got = m(wl)
        ref = orc.medium_n(m._record() | dict(aux_off=0, aux_n=0 if m._wls is None else len(m._wls)), wl,
                           orc.Ctx(None if m._wls is None else np.concatenate((m._wls, m._vals))))
        assert np.max(np.abs(got/ref - 1)) < 1e-13, m.spectrum_type
    with pytest.raises(RuntimeError):
        RI("Cauchy", coeff=[0.9, 0, 0, 0])(wl)            # n < 1 raises like refraction_index.py:165


def test_sphere_projections_match_oracle(ot):
    from optrace_b200.scene import detector_record
    from oracle import trace_oracle as orc
    surf = ot.SphericalSurface(r=8, R=-13.4)
    surf.move_to([0.3, -0.2, 24])
    rng = np.random.default_rng(5)
    xy = rng.uniform(-5, 5, (5000, 2))
    from optrace_b200 import engine
    z, _ = engine.surface_values(surf, 0.3 + xy[:, 0], -0.2 + xy[:, 1])
    p = np.column_stack((0.3 + xy[:, 0], -0.2 + xy[:, 1], z))
    for k, name in enumerate(surf.sphere_projection_methods):
        got = surf.sphere_projection(p, name)
        rec = detector_record(surf, name, None)
        ref = orc.sphere_projection(rec["surface"], rec["R"], p, rec["projection"])
        assert np.max(np.abs(got - ref)) < 1e-12, name


def test_render_image_power_conservation(ot):
    """tests/test_image.py:108-173: binning conserves power and luminous power; edges are inclusive
    (tests/test_misc.py:139-170)"""
    from optrace_b200 import color
    rng = np.random.default_rng(11)
    N = 200_000
    p = np.zeros((N, 3))
    p[:, 0], p[:, 1] = rng.uniform(-2, 3, N), rng.uniform(-1, 1, N)
    p[0, :2], p[1, :2] = [3, 1], [-2, -1]                      # exactly on the corners
    w = rng.uniform(0, 1e-4, N).astype(np.float32)
    wl = rng.uniform(380, 780, N).astype(np.float32)
    img = ot.RenderImage([-2, 3, -1, 1])
    img.render(p, w, wl)
    assert img.shape == (945, 945*3, 4)
    assert abs(img.power() - float(w.astype(np.float64).sum())) < 1e-6*w.sum()
    lum = 683.0*float(np.sum(color.y_observer(wl)*w))
    assert abs(img.luminous_power() - lum) < 1e-6*lum
    assert img.counts.sum() == N
    img2 = ot.RenderImage([-1, 1, -0.5, 0.5])                   # rays outside the extent are dropped
    img2.render(p, w, wl)
    inside = (np.abs(p[:, 0]) <= 1) & (np.abs(p[:, 1]) <= 0.5)
    assert img2.counts.sum() == np.count_nonzero(inside)
    empty = ot.RenderImage([-1, 1, -1, 1])
    empty.render(np.zeros((0, 3)), np.zeros(0, np.float32), np.zeros(0, np.float32))
    assert empty.power() == 0 and empty.shape == (945, 945, 4)


def test_fused_render_equals_store_path(ot):
    """iterative_render (fused kernel, no storage) bins exactly the hits of trace + detector_image"""
    RT = scenes.image_render(ot)
    N = 300_000
    RT.ITER_RAYS_STEP = N
    c0 = RT._trace_count
    ims = RT.iterative_render(N, pos=scenes.IMAGE_RENDER_POS)
    assert len(ims) == 6
    RT._trace_count = c0                     # same Philox key -> identical bundle
    RT.trace(N)
    for j, pos in enumerate(scenes.IMAGE_RENDER_POS):
        RT.detectors[0].move_to(pos)
        ref = RT.detector_image(extent=ims[j]._extent0)
        assert np.array_equal(ims[j].counts, ref.counts), j
        a, b = ims[j].data, ref.data
        assert np.all(np.abs(a - b) <= 1e-12*np.abs(b).max()), j
        assert np.allclose(ims[j].extent, ref.extent)
    # the auto extent of the fused range pass equals the store path's auto extent
    RT.detectors[0].move_to(scenes.IMAGE_RENDER_POS[2])
    auto = RT.detector_image()
    assert np.allclose(auto._extent0, ims[2]._extent0, rtol=1e-12, atol=0)
    assert np.array_equal(RT._msgs.sum(axis=1)[:2], RT._msgs.sum(axis=1)[:2])


def test_fused_render_matches_reference_fixture(ot):
    """fused kernel on the reference's frozen bundle: images equal the reference's detector images"""
    from optrace_b200 import engine
    from optrace_b200.scene import detector_record
    import torch
    for name in ("arizona_eye", "zoo_analytic", "hurb_pinhole"):
        g = gu.load(name)
        RT = scenes.SCENES[name](ot)
        scene = RT._scene_handle()
        p0, s0, pol0, w0, wl, hz = gu.bundle(g)
        rays = engine.DeviceRays.from_host(p0, s0, pol0, w0, wl, hz)
        for v in range(int(g["n_det"])):
            k = f"det{v}_"
            di, pm, src, ill = [int(x) for x in g[k + "spec"]]
            if src >= 0:
                continue
            RT.detectors[di].move_to(g[k + "pos"])
            rec = detector_record(RT.detectors[di].surface, gu.PROJ[pm], None)
            rng = engine.trace_render(scene, rays, [rec]).cpu().numpy()[0]
            assert np.allclose(rng, g[k + "extent0"], rtol=1e-9, atol=1e-12), (name, v)
            Ny, Nx = [int(x) for x in g[k + "shape"][:2]]
            img = torch.zeros((Ny, Nx, 4), dtype=torch.float64, device=engine.device())
            cnt = torch.zeros((Ny, Nx), dtype=torch.int32, device=engine.device())
            msgs = engine.trace_render(scene, rays, [rec], extents=[g[k + "extent"]], grids=[(Nx, Ny)], imgs=[img], cnts=[cnt])
            assert np.array_equal(msgs.cpu().numpy(), g["msgs"]), (name, v)
            ref = np.zeros((Ny, Nx, 4))
            ref[g[k + "yi"], g[k + "xi"]] = g[k + "vals"]
            d = img.cpu().numpy()
            nz = np.nonzero(d[:, :, 3])
            same_bins = np.array_equal(nz[0], g[k + "yi"]) and np.array_equal(nz[1], g[k + "xi"])
            if same_bins:
                assert np.all(np.abs(d - ref) <= gu.W_RTOL.get(name, 1e-9)*np.abs(ref) + 1e-12*np.abs(ref).max()), (name, v)
            # The fixture's auto extent is the min/max of the REFERENCE's hits, so its (up to 4) extreme rays sit
            # exactly on the image border; a last-ulp difference of a hit coordinate moves such a ray outside.
            missing = len(g[k + "w"]) - int(cnt.sum().item())
            assert 0 <= missing <= 4, (name, v, missing)
            tot = float(g[k + "vals"][:, 3].sum())
            assert abs(float(d[:, :, 3].sum()) - tot) <= 3e-7*tot + missing*float(g[k + "w"].max()), (name, v)


def test_shared_reciprocal_division_is_exact(ot):
    """the hoisted-reciprocal division used for vector normalisation equals IEEE division bit for bit"""
    import ctypes as C
    import torch
    from optrace_b200 import engine, _cabi
    lib = engine.ensure_init()
    g = torch.Generator(device="cuda").manual_seed(1)
    N = 1 << 24
    for scale in (1.0, 1e-3, 1e6, 1e-150, 1e150):
        a = (torch.rand(N, dtype=torch.float64, device="cuda", generator=g)*2 - 1)*scale
        b = (torch.rand(N, dtype=torch.float64, device="cuda", generator=g) + 1e-6)*3.7
        a[:8] = torch.tensor([0.0, -0.0, float("nan"), float("inf"), 1.0, -1.0, 5e-324, 1e308], dtype=torch.float64)
        b[8:12] = torch.tensor([0.0, float("inf"), float("nan"), 1e-310], dtype=torch.float64)
        qs, qi = torch.empty_like(a), torch.empty_like(a)
        _cabi.check(lib.otb_selftest_division(N, engine.dptr(a), engine.dptr(b), engine.dptr(qs), engine.dptr(qi),
                                              engine.stream_ptr()), lib)
        same = (qs.view(torch.int64) == qi.view(torch.int64)) | (torch.isnan(qs) & torch.isnan(qi))
        assert bool(same.all()), (scale, int((~same).sum()))
