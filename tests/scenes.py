"""Scene builders shared by the golden-vector generator (run against the REFERENCE: `ot` = optrace imported
from /root/reference) and by the tests / bench (run against this repo: `ot` = optrace_b200).  Because the
product is a drop-in for the reference's public API, the same construction code serves both.

C1..C5b follow the geometry sections of the reference's example scripts named in BASELINE.json `configs`
(SURVEY.md §8, table at the top); the zoo scenes exist to exercise every surface kind, medium model, filter
type and edge case (TIR, missed surfaces, outline hits, absorbed rays) of the hot path.
"""
import numpy as np


# ---- user callables for the function-surface scenes (module level so both engines see the same objects) ----
def cos_surface(x, y):
    return 0.1*np.cos(2*np.pi*x/2)


def parab_1d(r, a=0.02):
    return a*r**2 + 0.001*r**4


def parab_1d_deriv(r, a=0.02):
    return 2*a*r + 0.004*r**3


def parab_1d_mask(r):
    return r < 3.9


def saddle_2d(x, y):
    return 0.01*x**2 - 0.015*y**2 + 0.002*x*y


def saddle_2d_deriv(x, y):
    return 0.02*x + 0.002*y, -0.03*y + 0.002*x


def or_func_cone(x, y, f=40.0):
    """orientation function (ray_source.py:274-276): every ray aims at the point (0, 0, f) in front of its origin"""
    v = np.column_stack((-x, -y, np.ones_like(x)*f))
    return v/np.linalg.norm(v, axis=1)[:, np.newaxis]


def or_source(ot):
    """rectangular source with orientation="Function" in front of a detector (generator coverage, a18)"""
    RT = ot.Raytracer(outline=[-10, 10, -10, 10, -1, 30])
    RT.add(ot.RaySource(ot.RectangularSurface(dim=[4, 2]), orientation="Function", or_func=or_func_cone,
                        or_args=dict(f=25.0), pos=[0.5, -1, 0]))
    RT.add(ot.Detector(ot.RectangularSurface(dim=[8, 8]), pos=[0, 0, 20]))
    return RT


# ---- benchmark configs ---------------------------------------------------------------------------------
def spherical_aberration(ot):
    """C1: examples/spherical_aberration.py:18-64"""
    RT = ot.Raytracer(outline=[-10, 10, -10, 10, -25, 40])
    RT.add(ot.RaySource(ot.CircularSurface(r=1), divergence="None", spectrum=ot.presets.light_spectrum.d65,
                        pos=[0, 0, -15], s=[0, 0, 1]))
    RT.add(ot.RaySource(ot.RingSurface(r=4.5, ri=1), divergence="None", spectrum=ot.presets.light_spectrum.d65,
                        pos=[0, 0, -15], s=[0, 0, 1]))
    n = ot.RefractionIndex("Constant", n=1.5)
    RT.add(ot.Lens(ot.SphericalSurface(r=5, R=15), ot.SphericalSurface(r=5, R=-15), de=0.2, pos=[0, 0, 0], n=n))
    RT.add(ot.Detector(ot.RectangularSurface(dim=[20, 20]), pos=[0, 0, 23.]))
    return RT


def double_gauss(ot):
    """C2: examples/double_gauss.py:34-102 (Nikkor-Wakamiya 100 mm f/1.4, five point sources at -50 m)"""
    RT = ot.Raytracer(outline=[-2000, 2000, -22000, 2000, -50000, 180])
    g = 50000
    for deg in [0, 5, 10, 15, 20]:
        xp = g*np.tan(deg/180*np.pi)
        RT.add(ot.RaySource(ot.Point(), divergence="Isotropic", orientation="Converging", conv_pos=[0, 0, 0],
                            div_angle=0.03, pos=[0, -xp, -g], desc=f"{deg}°"))
    S = ot.SphericalSurface
    L_0 = ot.Lens(S(r=76/2, R=78.36), S(r=76/2, R=469.5), n=ot.RefractionIndex("Abbe", n=1.797, V=45.3),
                  pos=[0, 0, 0], d1=0, d2=9.8837)
    RT.add(L_0)
    L_1 = ot.Lens(S(r=64/2, R=50.3), S(r=62/2, R=74.38), n=ot.RefractionIndex("Abbe", n=1.773, V=49.4),
                  pos=[0, 0, L_0.back.pos[2]+0.1938], d1=0, d2=9.1085)
    RT.add(L_1)
    L_2 = ot.Lens(S(r=59/2, R=138.1), S(r=51/2, R=34.33), n=ot.RefractionIndex("Abbe", n=1.673, V=32.20),
                  pos=[0, 0, L_1.back.pos[2]+2.9457], d1=0, d2=2.3256)
    RT.add(L_2)
    RT.add(ot.Aperture(ot.RingSurface(ri=49.6/2, r=76/2), pos=[0, 0, L_2.back.pos[2]+16.07]))
    S_8 = S(r=57/2, R=-2907)
    L_3 = ot.Lens(S(r=48.8/2, R=-34.41), S_8, n=ot.RefractionIndex("Abbe", n=1.740, V=28.30),
                  pos=[0, 0, L_2.back.pos[2]+16.07+13], d1=0, d2=1.938)
    RT.add(L_3)
    L_4 = ot.Lens(S_8, S(r=60/2, R=-59.05), n=ot.RefractionIndex("Abbe", n=1.773, V=49.40),
                  pos=[0, 0, L_3.back.pos[2]+1e-6], d1=0, d2=12.403)
    RT.add(L_4)
    L_5 = ot.Lens(S(r=66.8/2, R=-150.9), S(r=67.8/2, R=-57.89), n=ot.RefractionIndex("Abbe", n=1.788, V=47.50),
                  pos=[0, 0, L_4.back.pos[2]+0.3876], d1=0, d2=8.333)
    RT.add(L_5)
    L_6 = ot.Lens(S(r=66/2, R=284.6), S(r=66/2, R=-253.2), n=ot.RefractionIndex("Abbe", n=1.788, V=47.50),
                  pos=[0, 0, L_5.back.pos[2]+0.1938], d1=0, d2=5.0388)
    RT.add(L_6)
    RT.add(ot.Detector(ot.RectangularSurface(dim=[86.53, 86.53]), pos=[0, 0, L_6.back.pos[2]+73.839]))
    return RT


def arizona_eye(ot, image=None):
    """C3: examples/arizona_eye_model.py:15-57 (chart at 0.6 m seen by the Arizona eye, retina detector)"""
    g, G_alpha, P = 0.6e3, 4, 4
    A = 1/g*1000
    G = g*np.tan(G_alpha/180*np.pi)
    OL = max(G, 8)
    RT = ot.Raytracer(outline=[-OL, OL, -OL, OL, -g, 28], no_pol=False)
    RSS = image if image is not None else ot.presets.image.ETDRS_chart_inverted([2*G, 2*G])
    sr_angle = np.rad2deg(np.arctan(1.4*P/2/g))
    RT.add(ot.RaySource(RSS, divergence="Isotropic", div_angle=sr_angle, pos=[0, 0, -g],
                        orientation="Converging", conv_pos=[0, 0, 0], desc="Chart"))
    RT.add(ot.presets.geometry.arizona_eye(adaptation=A, pupil=P))
    return RT


def image_render(ot, image=None):
    """C4: examples/image_render_many_rays.py:11-39 (no_pol, six detector positions)"""
    RSS = image if image is not None else ot.presets.image.tv_testcard2([4, 3])
    RT = ot.Raytracer(outline=[-8, 8, -8, 8, 0, 40], no_pol=True)
    div_angle = np.rad2deg(np.arctan(3/12)*1.2)
    RT.add(ot.RaySource(RSS, divergence="Isotropic", div_angle=div_angle, s=[0, 0, 1], pos=[0, 0, 0],
                        orientation="Converging", conv_pos=[0, 0, 12]))
    RT.add(ot.Lens(ot.SphericalSurface(r=3, R=8), ot.SphericalSurface(r=3, R=-8), de=0.1, pos=[0, 0, 12],
                   n=ot.RefractionIndex("Abbe", n=1.5, V=40)))
    RT.add(ot.Detector(ot.RectangularSurface(dim=[16, 16]), pos=[0, 0, 36]))
    return RT


IMAGE_RENDER_POS = [[0, 0, 15], [0, 0, 20], [0, 0, 25.], [0, 0, 29.], [0, 0, 31.], [0, 0, 36.]]


def cosine_surfaces(ot):
    """C5a: examples/cosine_surfaces.py:12-50"""
    RT = ot.Raytracer(outline=[-15, 15, -15, 15, 0, 80])
    RT.add(ot.RaySource(ot.CircularSurface(r=3), divergence="None", s=[0, 0, 1], pos=[0, 0, 0]))
    front = ot.FunctionSurface2D(func=cos_surface, r=5)
    back = front.copy()
    back.flip()
    back.rotate(90)
    L1 = ot.Lens(front, back, de=2, pos=[0, 0, 12], n=ot.presets.refraction_index.SF5)
    RT.add(L1)
    L2 = L1.copy()
    L2.move_to([0, 0, 18])
    RT.add(L2)
    RT.add(ot.IdealLens(r=9, D=50, pos=[0, 0, 40]))
    Det = ot.Detector(ot.RectangularSurface(dim=[14, 14]), pos=[0, 0, 24.4])
    RT.add(Det)
    Det2 = Det.copy()
    Det2.move_to([0, 0, 60])
    RT.add(Det2)
    return RT


def hurb_aperture(ot, name="Square"):
    """C5b: examples/hurb_apertures.py:12-56"""
    RT = ot.Raytracer(outline=[-5, 5, -5, 5, -1, 40], use_hurb=True, n0=ot.RefractionIndex("Constant", 1.33))
    if name == "Square":
        RT.add(ot.RaySource(ot.RectangularSurface(dim=[0.05, 0.05]), s=[0, 0, 1], pos=[0, 0, -1]))
        RT.add(ot.Aperture(ot.SlitSurface(dim=[2, 2], dimi=[0.05, 0.05]), pos=[0, 0, 0]))
    elif name == "Slit":
        RT.add(ot.RaySource(ot.RectangularSurface(dim=[0.05, 2]), s=[0, 0, 1], pos=[0, 0, -1]))
        RT.add(ot.Aperture(ot.SlitSurface(dim=[2.5, 2.5], dimi=[0.05, 2]), pos=[0, 0, 0]))
    elif name == "Edge":
        RT.add(ot.RaySource(ot.RectangularSurface(dim=[0.4, 1]), s=[0, 0, 1], pos=[0, 0.5, -1]))
        RT.add(ot.Aperture(ot.SlitSurface(dim=[2, 2], dimi=[1.8, 1.8]), pos=[0, 0.9, 0]))
    elif name == "Pinhole":
        RT.add(ot.RaySource(ot.CircularSurface(r=0.05), s=[0, 0, 1], pos=[0, 0, -1]))
        RT.add(ot.Aperture(ot.RingSurface(r=2, ri=0.025), pos=[0, 0, 0]))
    RT.add(ot.Detector(ot.RectangularSurface(dim=[1.5, 1.5]), pos=[0, 0, 30]))
    return RT


# ---- coverage scenes -------------------------------------------------------------------------------------
def zoo_analytic(ot):
    """conic / tilted / flat surfaces, every closed-form medium model, all filter types, ideal lens, n2 chain;
    wide bundles produce TIR, missed surfaces, outline hits"""
    RI = ot.RefractionIndex
    RT = ot.Raytracer(outline=[-12, 12, -12, 12, -20, 90], n0=RI("Cauchy", coeff=[1.0003, 1e-6, 0, 0]))
    RT.add(ot.RaySource(ot.CircularSurface(r=3), divergence="Isotropic", div_angle=10, pos=[0, 0, -15]))
    wls = np.linspace(380, 780, 21)
    T = ot.TransmissionSpectrum
    RT.add(ot.Filter(ot.CircularSurface(r=6), pos=[0, 0, -8],
                     spectrum=T("Data", wls=wls, vals=0.5 + 0.45*np.sin(wls/40.0))))
    RT.add(ot.Lens(ot.ConicSurface(r=6, R=18, k=-1.7), ot.ConicSurface(r=6, R=-22, k=0.4), d=3.2, pos=[0, 0.2, 0],
                   n=ot.presets.refraction_index.BK7, n2=RI("Constant", n=1.2)))
    rf = ot.RectangularSurface(dim=[9, 7])
    rf.rotate(25)
    RT.add(ot.Filter(rf, pos=[0.3, -0.2, 6], spectrum=T("Rectangle", wl0=420, wl1=690, val=0.8)))
    RT.add(ot.Lens(ot.TiltedSurface(r=5.5, normal=[0.15, -0.1, 1]), ot.TiltedSurface(r=5.5, normal_sph=[8, 140]),
                   d=2.5, pos=[0, 0, 12], n=RI("Sellmeier2", coeff=[1.045, 0.266, 0.206, 0.001, 0.3]),
                   n2=RI("Schott", coeff=[1.7, -0.01, 0.012, 0.0003, -1e-5, 1e-6])))
    RT.add(ot.Aperture(ot.RingSurface(r=7, ri=3.2), pos=[0, 0, 17]))
    RT.add(ot.Filter(ot.CircularSurface(r=7), pos=[0, 0, 19], spectrum=T("Constant", val=0.9, inverse=True)))
    RT.add(ot.Lens(ot.SphericalSurface(r=5, R=-9), ot.CircularSurface(r=5), d=1.0, pos=[0, 0, 24],
                   n=RI("Conrady", coeff=[1.9, 0.01, 0.0004]), n2=RI("Herzberger", coeff=[1.5, 0.004, 1e-4, -0.01, 1e-3, -1e-4])))
    sl = ot.SlitSurface(dim=[10, 10], dimi=[5, 3])
    sl.rotate(-15)
    RT.add(ot.Aperture(sl, pos=[0, 0, 30]))
    RT.add(ot.Filter(ot.CircularSurface(r=8), pos=[0, 0, 33], spectrum=T("Gaussian", mu=560, sig=60, val=0.95)))
    RT.add(ot.IdealLens(r=6, D=-14, pos=[0, 0, 38], n2=RI("Handbook of Optics 2", coeff=[1.6, 0.9, 0.02, 0.01])))
    RT.add(ot.Lens(ot.ConicSurface(r=7, R=25, k=-1.0), ot.SphericalSurface(r=7, R=-40), d=2.0, pos=[0, 0, 46],
                   n=RI("Sellmeier3", coeff=[5.684027565E-1, 5.101829712E-3, 1.726177391E-1, 1.821153936E-2,
                                            2.086189578E-2, 2.620722293E-2, 1.130748688E-1, 1.069792721E1]),
                   n2=RI("Data", wls=np.linspace(380, 780, 9), vals=1.3 + 0.05*np.cos(np.linspace(0, 2, 9)))))
    RT.add(ot.IdealLens(r=9, D=30, pos=[0, 0, 55]))
    RT.add(ot.Detector(ot.RectangularSurface(dim=[16, 16]), pos=[0, 0, 70]))
    RT.add(ot.Detector(ot.SphericalSurface(r=9, R=-20), pos=[0, 0, 62]))
    RT.add(ot.Detector(ot.TiltedSurface(r=9, normal=[0.1, 0.2, 1]), pos=[0, 0, 10]))
    RT.add(ot.Detector(ot.CircularSurface(r=4), pos=[0, 0, 3]))
    return RT


def zoo_numeric(ot):
    """asphere, function (1-D with derivative+mask, 2-D with and without derivative, flipped/rotated) and data
    (1-D, 2-D) surfaces: numeric Illinois hit finding, finite-difference and analytic normals"""
    RI = ot.RefractionIndex
    RT = ot.Raytracer(outline=[-12, 12, -12, 12, -20, 80])
    RT.add(ot.RaySource(ot.CircularSurface(r=2.5), divergence="Lambertian", div_angle=4, pos=[0, 0, -15]))
    RT.add(ot.Lens(ot.AsphericSurface(r=4, R=12, k=-0.6, coeff=[1e-3, -2e-5, 1e-7]),
                   ot.AsphericSurface(r=4, R=-15, k=0.3, coeff=[-5e-4, 1e-5]), d=2.5, pos=[0, 0, 0],
                   n=RI("Sellmeier1", coeff=[1.03961212, 0.00600069867, 0.231792344, 0.0200179144, 1.01046945, 103.560653])))
    f1 = ot.FunctionSurface1D(r=4, func=parab_1d, deriv_func=parab_1d_deriv, mask_func=parab_1d_mask,
                              func_args=dict(a=0.025), deriv_args=dict(a=0.025), parax_roc=20)
    b1 = f1.copy()
    b1.flip()
    RT.add(ot.Lens(f1, b1, d=2.0, pos=[0, 0, 8], n=RI("Abbe", n=1.6, V=35)))
    f2 = ot.FunctionSurface2D(r=4.5, func=saddle_2d, deriv_func=saddle_2d_deriv)
    f2.rotate(30)
    b2 = ot.FunctionSurface2D(r=4.5, func=saddle_2d)
    b2.flip()
    b2.rotate(-40)
    RT.add(ot.Lens(f2, b2, d=1.8, pos=[0, 0, 15], n=RI("Constant", n=1.45)))
    Y, X = np.mgrid[-5:5:120j, -5:5:120j]
    d2 = ot.DataSurface2D(r=5, data=0.012*(X**2 + 0.7*Y**2) + 0.003*X*Y)
    d2.rotate(20)
    r1 = np.linspace(0, 5, 150)
    d1 = ot.DataSurface1D(r=5, data=-0.015*r1**2 + 2e-4*r1**4)
    RT.add(ot.Lens(d2, d1, d=2.2, pos=[0, 0, 23], n=RI("Cauchy", coeff=[1.52, 0.004, 1e-5, 0])))
    d3 = ot.DataSurface2D(r=5, data=0.01*(X**2 + Y**2))
    d3.flip()
    RT.add(ot.Lens(ot.CircularSurface(r=5), d3, d=1.5, pos=[0, 0, 30], n=RI("Constant", n=1.7)))
    RT.add(ot.Detector(ot.RectangularSurface(dim=[14, 14]), pos=[0, 0, 45]))
    return RT


def microscope(ot, no_pol=False):
    """the reference's own benchmark scene (tests/benchmark.py:16-66): 60x microscope objective + tube lens (.zmx),
    eyepiece (.zmx), Arizona eye model, cell image source — 57 tracing surfaces, glasses from four .agf catalogues.
    The resource files travel with the vendored reference (oracle/_ref, tools/vendor_reference.py); the element
    positions the reference derives with its paraxial analysis (TMA, out of scope here) come from the fixture
    tests/golden/load_zmx.json (tools/gen_golden_load.py)."""
    import json
    import pathlib
    root = pathlib.Path(__file__).resolve().parent.parent
    res = root / "oracle" / "_ref" / "examples" / "resources"
    if not res.exists():
        res = pathlib.Path("/root/reference/examples/resources")
    pos = json.loads((root / "tests" / "golden" / "load_zmx.json").read_text())["benchmark"]
    RT = ot.Raytracer(outline=[-50, 50, -50, 50, -30, 430], no_pol=no_pol)
    RT.add(ot.RaySource(ot.presets.image.cell([100e-3, 100e-3]), divergence="Lambertian", pos=[0, 0, -0.00000001],
                        s=[0, 0, 1], div_angle=50, desc="Cell"))
    n_dict = {}
    for f in ("schott", "ohara", "hikari", "hoya"):
        n_dict |= ot.load_agf(str(res / "materials" / f"{f}.agf"))
    G = ot.load_zmx(str(res / "microscope" / "Nikon_1p25NA_60x_US7889433B2_MultiConfig_v2.zmx"), n_dict=n_dict)
    RT.n0 = G.n0
    RT.add(ot.Group(G.lenses[:18]))
    tube = ot.Group(G.lenses[20:24])
    tube.move_to(pos["tube_pos"])
    RT.add(tube)
    eyepiece = ot.load_zmx(str(res / "eyepiece" / "UK565851-1.zmx"), n_dict=n_dict)
    eyepiece.remove(eyepiece.detectors)
    eyepiece.move_to(pos["eyepiece_pos"])
    RT.add(eyepiece)
    eye = ot.presets.geometry.arizona_eye()
    eye.move_to(pos["eye_pos"])
    RT.add(eye)
    return RT


def microscope_available() -> bool:
    import pathlib
    root = pathlib.Path(__file__).resolve().parent.parent
    return (root / "oracle" / "_ref" / "examples" / "resources").exists() or pathlib.Path("/root/reference/examples/resources").exists()


SCENES = dict(spherical_aberration=spherical_aberration, double_gauss=double_gauss, arizona_eye=arizona_eye,
              image_render=image_render, cosine_surfaces=cosine_surfaces,
              hurb_square=lambda ot: hurb_aperture(ot, "Square"), hurb_pinhole=lambda ot: hurb_aperture(ot, "Pinhole"),
              hurb_edge=lambda ot: hurb_aperture(ot, "Edge"), hurb_slit=lambda ot: hurb_aperture(ot, "Slit"),
              zoo_analytic=zoo_analytic, zoo_numeric=zoo_numeric)


# the reference's benchmark scene needs the .zmx / .agf files that travel with the vendored reference
if microscope_available():
    SCENES["microscope"] = microscope


# ---- known-answer surfaces of the reference's own tests (tests/test_surface.py:158-174, 204-219) -----------
def kat_f2d(x, y):
    return x**2 + y**2/2


def kat_f1d_sq(r):
    return r**2


def kat_f1d_lin(r):
    return r + 0.1*r**2


def kat_surfaces(ot):
    """(surface, normal at (x0+1, y0+0.5) or None, values at [(x0+1, y0+0.5), (x0, y0-0.5)] - z0 or None)"""
    return [
        (ot.SphericalSurface(r=5, R=-10), [0.1, 0.05, 0.99373035], [-0.06269654, -0.01250782]),
        (ot.ConicSurface(r=5, R=12, k=3), [-0.08444007, -0.04222003, 0.9955337], [0.05254347, 0.01043481]),
        (ot.ConicSurface(r=5, R=-12, k=3), [0.08444007, 0.04222003, 0.9955337], None),
        (ot.AsphericSurface(r=5, R=12, k=3, coeff=[0, 1e-4, 1e-8]), [-0.08493345, -0.04246673, 0.99548123],
         [0.05269974, 0.01044106]),
        (ot.TiltedSurface(r=2, normal_sph=[20, 50]), [0.21984631, 0.26200263, 0.93969262], [-0.37336424, 0.13940869]),
        (ot.DataSurface1D(r=3, data=-1+np.linspace(0, 3, 200)**1.25), [-0.70594412, -0.35297206, 0.61404693],
         [1.14965824, 0.42044821]),
        (ot.FunctionSurface2D(r=4, func=kat_f2d), [-0.87287156, -0.21821789, 0.43643578], [1.125, 0.125]),
        (ot.DataSurface2D(r=3, data=1+np.mgrid[-3:3:100j, -3:3:100j][1] + np.linspace(-3, 3, 100)**2),
         [0, -0.894427191, 0.447213595], [0.75, -0.25]),
        (ot.FunctionSurface1D(r=4, func=kat_f1d_sq), None, None),
        (ot.FunctionSurface1D(r=4, func=kat_f1d_lin), None, [1.24303399, 0.525]),
    ]


def hit_test_surfaces(ot):
    """one instance of every surface kind for the hit-finding property test (tests/test_surface.py:237-268)"""
    rs = ot.RectangularSurface(dim=[3, 2])
    rs.rotate(30)
    sl = ot.SlitSurface(dim=[4, 3], dimi=[1, 0.5])
    return [ot.CircularSurface(r=3), rs, ot.RingSurface(r=3, ri=0.7), sl, ot.SphericalSurface(r=3, R=5),
            ot.SphericalSurface(r=3, R=-7), ot.ConicSurface(r=3, R=-8, k=-2.5), ot.ConicSurface(r=2.5, R=4, k=0.9),
            ot.TiltedSurface(r=3, normal=[0.2, -0.1, 1]), ot.AsphericSurface(r=3, R=9, k=-0.4, coeff=[1e-3, 2e-5]),
            ot.DataSurface1D(r=3, data=0.03*np.linspace(0, 3, 120)**2),
            ot.DataSurface2D(r=3, data=0.02*(np.mgrid[-3:3:90j, -3:3:90j][1]**2) + 0.01*np.mgrid[-3:3:90j, -3:3:90j][0]),
            ot.FunctionSurface2D(r=4, func=kat_f2d, z_min=0, z_max=16), ot.FunctionSurface1D(r=4, func=kat_f1d_lin)]
