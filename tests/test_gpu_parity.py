"""GPU parity tests (run with -m gpu on a B200): the CUDA engine, called through the drop-in Python API and
the C ABI behind it, against (a) the golden fixtures = outputs of the reference itself on frozen bundles and
(b) the pinned CPU oracle on larger seeded bundles.

Tolerances (BASELINE.json north_star): positions, directions, weights, polarisation within 1e-9 relative;
message counters exact; detector pixel COUNTS exact (rays on bin edges may move by one bin: at most +-1 per
bin), XYZW sums within 1e-9 relative (atomic accumulation order differs from np.add.at)."""
import numpy as np
import pytest

import golden_util as gu
import scenes

pytestmark = pytest.mark.gpu

RTOL = 1e-9
# scenes without transcendental functions on the ray path: every float64 operation is reproduced bit for bit
BIT_EXACT = {"double_gauss", "spherical_aberration", "image_render", "hurb_square", "hurb_edge", "hurb_slit"}
NEEDS_USERFUNC = {"cosine_surfaces", "zoo_numeric"}
NAMES = list(scenes.SCENES)


@pytest.fixture(scope="module")
def ot():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import optrace_b200 as ot
    from optrace_b200 import engine
    engine.ensure_init()
    return ot


def _trace_fixture(ot, name, arithmetic=None):
    g = gu.load(name)
    RT = scenes.SCENES[name](ot)
    if arithmetic is not None:
        RT.arithmetic = arithmetic
    p0, s0, pol0, w0, wl, hz = gu.bundle(g)
    RT.trace_rays(p0, s0, pol0, w0, wl, hurb_z=hz, N_list=g["N_list"])
    return RT, g


@pytest.mark.parametrize("arithmetic", ["relaxed", "exact"])
@pytest.mark.parametrize("name", NAMES)
def test_trace_matches_reference(ot, name, arithmetic):
    """both floating-point contracts of the lens step (Raytracer.arithmetic) against the reference's output"""
    RT, g = _trace_fixture(ot, name, arithmetic)
    R = RT.rays
    assert R.p_list.shape == g["p_list"].shape and R.p_list.dtype == np.float64 and R.p_list.flags.f_contiguous
    assert R.w_list.dtype == np.float32 and R.n_list.dtype == np.float64 and R.wl_list.dtype == np.float32
    assert np.array_equal(RT._msgs, g["msgs"]), (RT._msgs, g["msgs"])
    errs = dict(p=gu.vecrel(R.p_list, g["p_list"]), s=gu.vecrel(R.s0_list, g["s_list"]),
                w=gu.maxrel(R.w_list, g["w_list"]), n=gu.maxrel(R.n_list, g["n_list"]))
    if "pol_list" in g:
        assert R.pol_list.dtype == np.float32
        errs["pol"] = gu.vecrel(R.pol_list, g["pol_list"])
    else:
        assert np.all(np.isnan(R.pol_list))
    print(name, arithmetic, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        tol = gu.W_RTOL.get(name, RTOL) if k == "w" else RTOL
        if arithmetic == "relaxed":
            # opt-in mode, not the drop-in contract: float64 quantities inherit the conditioning of the reference's
            # own formulas (sphere intersection from 50 m away: 7 digits cancel), float32 storage flips by one ulp
            tol = max(tol, 2e-7) if k in ("w", "pol") else 2e-8
        assert v <= tol, (k, v)
    assert np.array_equal(R.wl_list, g["wl"])


@pytest.mark.parametrize("name", NAMES)
def test_detector_images_match_reference(ot, name):
    RT, g = _trace_fixture(ot, name)
    for v in range(int(g["n_det"])):
        k = f"det{v}_"
        di, pm, src, ill = [int(x) for x in g[k + "spec"]]
        RT.detectors[di].move_to(g[k + "pos"])
        ext = g.get(k + "user_extent")
        img = RT.detector_image(di, None if src < 0 else src, extent=ext, projection_method=gu.PROJ[pm])
        assert img.shape == tuple(g[k + "shape"])
        # auto extents are min/max of hit coordinates: same tolerance as the positions themselves
        assert np.allclose(img.extent, g[k + "extent"], rtol=RTOL, atol=1e-12)
        assert np.allclose(img._extent0, g[k + "extent0"], rtol=RTOL, atol=1e-12)
        _check_image(name, v, g, k, img)


def _check_image(name, v, g, k, img):
    """pixel counts exact except for hits on bin edges, XYZW values always compared on the bins whose count agrees"""
    data, cnt = img.data, img.counts
    ref = np.zeros(data.shape)
    ref[g[k + "yi"], g[k + "xi"]] = g[k + "vals"]
    # reference counts: the fixture hits binned by the reference's own misc.binning_indices_2d (tools/gen_golden.py)
    refcnt = np.zeros(cnt.shape, dtype=np.int64)
    refcnt[g[k + "cyi"], g[k + "cxi"]] = g[k + "cnt"]
    diff = cnt.astype(np.int64) - refcnt
    n_edge = int(g[k + "n_edge"])        # hits within 1e-6 of a bin edge (in bin units): only those may move by one bin
    assert np.abs(diff).max() <= 1 and np.count_nonzero(diff) <= 2*n_edge, (name, v, np.count_nonzero(diff), n_edge)
    assert abs(int(cnt.sum()) - int(refcnt.sum())) <= n_edge, (name, v, cnt.sum(), refcnt.sum())
    same = diff == 0
    assert np.count_nonzero(same & (refcnt > 0)) >= np.count_nonzero(refcnt > 0) - 2*n_edge
    scale = np.abs(ref).max(axis=(0, 1))
    tol = gu.W_RTOL.get(name, RTOL)
    ok = np.abs(data - ref) <= tol*np.abs(ref) + 1e-12*scale
    assert np.all(ok[same]), (name, v, float(np.max(np.abs(data - ref)[same])))
    tot = float(np.sum(g[k + "vals"][:, 3]))
    if not np.count_nonzero(diff):
        assert abs(img.power() - tot) <= tol*max(1e-30, tot)


@pytest.mark.parametrize("name", NAMES)
def test_large_fixtures_match_reference(ot, name):
    """the reference's output on frozen 1e5-ray bundles (tests/golden_large/, written by
    __graft_entry__.build() with tools/gen_golden.py in the container that holds the reference; git-ignored, travels
    with the snapshot): ray storage within 1e-9, messages exact, every detector image of the fixture"""
    path = gu.ROOT / "tests" / "golden_large" / f"{name}.npz"
    if not path.exists():
        pytest.skip("large fixtures not generated (no reference in the build container)")
    g = dict(np.load(path))
    RT = scenes.SCENES[name](ot)
    p0, s0, pol0, w0, wl, hz = gu.bundle(g)
    assert p0.shape[0] >= 10_000
    RT.trace_rays(p0, s0, pol0, w0, wl, hurb_z=hz, N_list=g["N_list"])
    R = RT.rays
    assert np.array_equal(RT._msgs, g["msgs"]), (RT._msgs, g["msgs"])
    # the index array travels as its per-section column sums only (snapshot size); element-wise it is compared on the
    # small fixtures and against the oracle at 2e5 rays
    exact = name in BIT_EXACT          # closed-form scenes: float32-stored quantities must agree without exception
    fw = (lambda a, b: a) if exact else gu.f32_flips
    fww = (lambda a, b: a) if (exact or name in gu.W_RTOL) else gu.f32_flips     # W_RTOL scenes: float32 exp, see there
    errs = dict(p=gu.vecrel(R.p_list, g["p_list"]), s=gu.vecrel(R.s0_list, g["s_list"]),
                w=gu.maxrel(fww(R.w_list, g["w_list"]), g["w_list"]), n=gu.maxrel(R.n_list.sum(axis=0), g["n_list_sum"]))
    if "pol_list" in g:
        errs["pol"] = gu.vecrel(fw(R.pol_list, g["pol_list"]), g["pol_list"])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v <= (gu.W_RTOL.get(name, RTOL) if k == "w" else RTOL), (k, v)
    for v in range(int(g["n_det"])):
        k = f"det{v}_"
        di, pm, src, ill = [int(x) for x in g[k + "spec"]]
        RT.detectors[di].move_to(g[k + "pos"])
        img = RT.detector_image(di, None if src < 0 else src, extent=g.get(k + "user_extent"), projection_method=gu.PROJ[pm])
        assert img.shape == tuple(g[k + "shape"])
        assert np.allclose(img.extent, g[k + "extent"], rtol=RTOL, atol=1e-12)
        _check_image(name, v, g, k, img)
    # spectra rendered by the reference (float32 accumulation there, float64 here)
    for key, spec in (("spec_det0", RT.detector_spectrum(0)), ("spec_src0", RT.source_spectrum(0))):
        assert np.array_equal(spec._wls, g[key + "_wls"])
        assert np.allclose(spec._vals, g[key + "_vals"], rtol=2e-5, atol=1e-9*np.max(g[key + "_vals"]))


@pytest.mark.parametrize("name", NAMES)
def test_device_generation_and_oracle_parity(ot, name):
    """device-generated bundle (Philox) traced on the GPU vs the pinned oracle on the same bundle, 200k rays (100k
    for the scenes whose numeric surfaces make the numpy oracle slow), every scene incl. the HURB ones: the normal
    deviates the kernel drew are replayed for the oracle (otb_hurb_normals)"""
    from optrace_b200.scene import flatten_raytracer
    from optrace_b200 import engine
    from oracle import trace_oracle as orc
    import torch
    RT = scenes.SCENES[name](ot)
    N = 100_000 if name in NEEDS_USERFUNC else 200_000
    RT.trace(N)
    R = RT.rays
    # initial bundle = section 0 of the store + regenerate directions through the generator
    N_list = R.N_list
    rays = RT._generate(N_list, 0, N, (int(RT.seed) << 20) + RT._trace_count)
    p0 = rays.p0.cpu().numpy().reshape((N, 3), order="F")
    s0 = rays.s0.cpu().numpy().reshape((N, 3), order="F")
    pol0 = None if RT.no_pol else rays.pol0.cpu().numpy().reshape((N, 3), order="F")
    w0, wl = rays.w0.cpu().numpy(), rays.wl.cpu().numpy()
    assert np.array_equal(p0, R.p_list[:, 0]) and np.array_equal(wl, R.wl_list)
    assert np.all(s0[:, 2] > 0) and np.allclose(np.linalg.norm(s0, axis=1), 1, atol=1e-12)
    if pol0 is not None:
        assert np.max(np.abs(np.sum(pol0.astype(np.float64)*s0, axis=1))) < 1e-6
    assert abs(float(w0.astype(np.float64).sum()) - sum(rs.power for rs in RT.ray_sources)) < 1e-3
    fs = flatten_raytracer(RT)
    hz = None
    if fs.n_hurb:
        hz = engine.hurb_normals(RT._scene.lib, N, (int(RT.seed) << 20) + RT._trace_count, 0, fs.n_hurb)
        assert abs(hz.mean()) < 0.01 and abs(hz.std() - 1) < 0.01
    ref = orc.trace(fs, p0, s0, pol0, w0, wl, hz)
    assert np.array_equal(RT._msgs, ref["msgs"])
    for a, b, nm in ((R.p_list, ref["p"], "p"), (R.s0_list, ref["s"], "s")):
        e = gu.vecrel(a, b)
        assert e <= RTOL, (nm, e)
    fw = (lambda a, b: a) if name in BIT_EXACT else gu.f32_flips
    fww = (lambda a, b: a) if (name in BIT_EXACT or name in gu.W_RTOL) else gu.f32_flips
    for a, b, nm in ((fww(R.w_list, ref["w"]), ref["w"], "w"), (R.n_list, ref["n"], "n")):
        e = gu.maxrel(a, b)
        assert e <= (gu.W_RTOL.get(name, RTOL) if nm == "w" else RTOL), (nm, e)
    if not RT.no_pol:
        assert gu.vecrel(fw(R.pol_list, ref["pol"]), ref["pol"]) <= RTOL


@pytest.mark.parametrize("name", ["double_gauss", "arizona_eye", "zoo_analytic", "image_render"])
def test_specialised_kernels_are_bit_identical(ot, name):
    """Raytracer.compile(): scene-specialised kernels give exactly the generic kernels' results (store path, fused
    render path) and keep matching the reference fixture"""
    g = gu.load(name)
    p0, s0, pol0, w0, wl, hz = gu.bundle(g)
    res = []
    for spec in (False, True):
        RT = scenes.SCENES[name](ot)
        RT.use_specialised_kernels = spec
        if spec:
            assert RT.compile() and RT._scene.specialised
        RT.trace_rays(p0, s0, pol0, w0, wl, hurb_z=hz, N_list=g["N_list"])
        assert RT._scene.specialised == spec
        R = RT.rays
        img = RT.detector_image()
        RT.ITER_RAYS_STEP = 50_000
        RT._trace_count = 0
        fused = RT.iterative_render(100_000)[0]
        res.append((R.p_list.copy(), R.s0_list.copy(), R.w_list.copy(), R.n_list.copy(),
                    None if RT.no_pol else R.pol_list.copy(), RT._msgs.copy(), img.data, fused.data, fused.counts))
    for k, (a, b) in enumerate(zip(*res)):
        if a is None:
            continue
        if k in (6, 7):      # images: atomic accumulation order is not deterministic, values agree to rounding
            assert np.allclose(a, b, rtol=1e-13, atol=0)
        else:
            assert np.array_equal(a, b, equal_nan=True), k
    assert gu.vecrel(res[1][0], g["p_list"]) <= RTOL and np.array_equal(res[1][5][:, :0], res[1][5][:, :0])


@pytest.mark.parametrize("name", ["double_gauss", "spherical_aberration", "image_render"])
def test_exact_arithmetic_is_bit_identical_on_closed_form_scenes(ot, name):
    """Raytracer.arithmetic = "exact": spherical / flat scenes without transcendental functions reproduce the
    reference's float64 results bit for bit (positions, directions, weights, indices, polarisation, messages)"""
    RT, g = _trace_fixture(ot, name, "exact")
    R = RT.rays
    assert np.array_equal(R.p_list, g["p_list"], equal_nan=True)
    assert np.array_equal(R.s0_list, g["s_list"], equal_nan=True)
    assert np.array_equal(R.w_list, g["w_list"]) and np.array_equal(R.n_list, g["n_list"])
    if "pol_list" in g:
        assert np.array_equal(R.pol_list, g["pol_list"], equal_nan=True)
    assert np.array_equal(RT._msgs, g["msgs"])


def test_relaxed_arithmetic_error_level(ot):
    """the opt-in relaxed contract: every operation is accurate to 1-2 ulp; what differs from the reference is the
    rounding of ill-conditioned expressions of the reference itself (documented in Raytracer.arithmetic)"""
    RT, g = _trace_fixture(ot, "spherical_aberration", "relaxed")       # sources at 3 mm: well conditioned
    R = RT.rays
    assert gu.vecrel(R.p_list, g["p_list"]) < 1e-12 and gu.vecrel(R.s0_list, g["s_list"]) < 1e-12
    assert np.array_equal(RT._msgs, g["msgs"])
    RT, g = _trace_fixture(ot, "double_gauss", "relaxed")               # sources at 50 m
    assert gu.vecrel(RT.rays.p_list, g["p_list"]) < 5e-9 and np.array_equal(RT._msgs, g["msgs"])


def test_deferred_status_collects_messages_at_the_detector_call(ot):
    """Raytracer.deferred_status (extension): trace() returns without the host synchronisation; the message counters
    and the device status arrive with the next synchronising call and equal those of an ordinary trace; a device-side
    error is raised there instead of inside trace()"""
    g = gu.load("double_gauss")
    p0, s0, pol0, w0, wl, hz = gu.bundle(g)
    RT = scenes.SCENES["double_gauss"](ot)
    RT.trace_rays(p0, s0, pol0, w0, wl, N_list=g["N_list"])
    msgs_ref = RT._msgs.copy()
    assert msgs_ref.any()
    img_ref = RT.detector_image()
    RT2 = scenes.SCENES["double_gauss"](ot)
    RT2.deferred_status = True
    RT2.trace_rays(p0, s0, pol0, w0, wl, N_list=g["N_list"])
    assert RT2.__dict__.get("_pending_trace") is not None          # nothing read back yet
    img = RT2.detector_image()
    assert RT2.__dict__.get("_pending_trace") is None
    assert np.array_equal(RT2._msgs, msgs_ref)
    assert np.array_equal(img.counts, img_ref.counts)
    # a device-side error (generated direction with s_z <= 0, ray_source.py:353) is raised by finish_trace()
    RT3 = ot.Raytracer(outline=[-10, 10, -10, 10, -10, 10])
    RT3.add(ot.RaySource(ot.Point(), divergence="Isotropic", div_angle=80, s=[1, 0, 0.2]))
    RT3.deferred_status = True
    RT3.trace(10000)                   # returns
    assert RT3.__dict__.get("_pending_trace") is not None
    with pytest.raises(RuntimeError):
        RT3.finish_trace()


def test_ray_store_is_recycled_only_when_unobservable(ot):
    """a repeated trace overwrites the previous device planes in place (no 8 GB allocator round trip) unless the
    caller still holds the previous RayStorage: that one keeps its data"""
    g = gu.load("double_gauss")
    p0, s0, pol0, w0, wl, hz = gu.bundle(g)
    RT = scenes.SCENES["double_gauss"](ot)
    RT.trace_rays(p0, s0, pol0, w0, wl, N_list=g["N_list"])
    ptr1 = RT.rays._dev.p.data_ptr()
    ref_p = RT.rays.p_list.copy()
    RT.trace_rays(p0, s0, pol0, w0, wl, N_list=g["N_list"])
    assert RT.rays._dev.p.data_ptr() == ptr1                      # recycled: nobody else could see the old storage
    assert np.array_equal(RT.rays.p_list, ref_p)
    held = RT.rays                                                   # the caller keeps the storage of this trace ...
    RT.trace_rays(p0 + np.array([1e-3, 0, 0]), s0, pol0, w0, wl, N_list=g["N_list"])
    assert RT.rays._dev.p.data_ptr() != held._dev.p.data_ptr()     # ... so the next trace got planes of its own
    assert not np.array_equal(RT.rays.p_list, ref_p)
    assert np.array_equal(held.p_list, ref_p)                       # and the held storage still has its data
