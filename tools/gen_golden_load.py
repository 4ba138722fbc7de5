"""Fixture for the ZEMAX importers and the collision check: what the REFERENCE builds from the files of its own
benchmark (tests/benchmark.py: microscope objective + tube + eyepiece .zmx, four .agf catalogues), and its
check_collision verdicts on a few surface pairs.  Also records the element positions of the assembled benchmark
scene, which the reference derives with its paraxial analysis (TMA: out of this repository's scope).
Run in the development container only:  python tools/gen_golden_load.py  ->  tests/golden/load_zmx.json"""
import json
import pathlib
import sys
import warnings

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
from refharness import import_reference  # noqa: E402

ot = import_reference()
RES = pathlib.Path("/root/reference/examples/resources")
warnings.simplefilter("ignore")
ot.global_options.show_warnings = False


def medium(n):
    if n is None:
        return None
    return dict(type=n.spectrum_type, coeff=None if n.coeff is None else [float(v) for v in n.coeff],
                val=float(n.val), V=None if n.V is None else float(n.V))


def surface(s):
    d = dict(cls=type(s).__name__, r=float(s.r), pos=[float(v) for v in s.pos], z_min=float(s.z_min), z_max=float(s.z_max))
    for k in ("R", "k", "ri"):
        if hasattr(s, k):
            d[k] = float(getattr(s, k))
    if hasattr(s, "coeff"):
        d["coeff"] = [float(v) for v in s.coeff]
    if hasattr(s, "dim"):
        d["dim"] = [float(v) for v in s.dim]
    return d


def group(G):
    out = dict(n0=medium(G.n0), long_desc=G.long_desc, elements=[])
    for el in G.elements:
        e = dict(cls=type(el).__name__, desc=el.desc, pos=[float(v) for v in el.pos], front=surface(el.front))
        if el.has_back():
            e.update(back=surface(el.back), d1=float(el.d1), d2=float(el.d2))
        if hasattr(el, "n"):
            e.update(n=medium(el.n), n2=medium(el.n2))
        out["elements"].append(e)
    return out


n_dict = {}
cat = {}
for f in ("schott", "ohara", "hikari", "hoya"):
    d = ot.load_agf(str(RES / "materials" / f"{f}.agf"))
    cat[f] = {k: medium(v) for k, v in d.items()}
    n_dict |= d
G = ot.load_zmx(str(RES / "microscope" / "Nikon_1p25NA_60x_US7889433B2_MultiConfig_v2.zmx"), n_dict=n_dict)
E = ot.load_zmx(str(RES / "eyepiece" / "UK565851-1.zmx"), n_dict=n_dict)
out = dict(catalogues=cat, microscope=group(G), eyepiece=group(E))

# the assembled benchmark scene (tests/benchmark.py:16-66): positions from the reference's paraxial analysis
RT = ot.Raytracer(outline=[-50, 50, -50, 50, -30, 430])
RSS = ot.presets.image.cell([100e-3, 100e-3])
RS = ot.RaySource(RSS, divergence="Lambertian", pos=[0, 0, -0.00000001], s=[0, 0, 1], div_angle=50, desc="Cell")
RT.add(RS)
RT.n0 = G.n0
objective = ot.Group(G.lenses[:18])
RT.add(objective)
tube = ot.Group(G.lenses[20:24])
tube.move_to(G.lenses[20].pos - [0, 0, 150])
RT.add(tube)
E.remove(E.detectors)
tma = ot.TMA(objective.lenses + tube.lenses, n0=G.n0)
z_img0 = tma.image_position(RS.pos[2])
eyep_f0 = E.tma().focal_points[0]
eyep_pos = [0, 0, E.lenses[0].pos[2] - (eyep_f0 - z_img0)]
E.move_to(eyep_pos)
RT.add(E)
eye = ot.presets.geometry.arizona_eye()
exit_pupil = RT.tma().pupil_position(0.38)[1]
entrance_pupil_eye = eye.tma().pupil_position(eye.apertures[0].pos[2])[0]
eye_pos = [0, 0, exit_pupil + (eye.pos[2] - entrance_pupil_eye)]
eye.move_to(eye_pos)
RT.add(eye)
out["benchmark"] = dict(tube_pos=[float(v) for v in tube.pos], eyepiece_pos=[float(v) for v in eyep_pos],
                        eye_pos=[float(v) for v in eye_pos], n_surfaces=len(RT.tracing_surfaces),
                        surfaces_z=[float(s.pos[2]) for s in RT.tracing_surfaces])

# check_collision known answers (raytracer.py:581-664)
S = ot.SphericalSurface
cases = []


def cc(name, a, b):
    c, x, y, z = ot.Raytracer.check_collision(a, b)
    cases.append(dict(name=name, coll=bool(c), n=int(x.shape[0]), first=[float(x[0]), float(y[0]), float(z[0])] if x.shape[0] else None))


a = S(r=3, R=5); a.move_to([0, 0, 0])
b = S(r=3, R=-5); b.move_to([0, 0, 1.2])
cc("biconvex_thin", a, b)
b2 = S(r=3, R=-5); b2.move_to([0, 0, 2.5])
cc("biconvex_ok", a, b2)
c1 = ot.CircularSurface(r=2); c1.move_to([0.5, 0, 0.3])
cc("sphere_vs_circle", a, c1)
t = ot.TiltedSurface(r=3, normal=[0.4, 0, 1]); t.move_to([0, 0, 0.5])
cc("sphere_vs_tilted", a, t)
p = ot.Point(); p.move_to([1.0, 0.5, 0.05])
cc("point_front", p, a)
cc("surface_point", a, p)
ln = ot.Line(r=2.5, angle=30); ln.move_to([0, 0, 0.4])
cc("line_front", ln, a)
r1 = ot.RectangularSurface(dim=[2, 2]); r1.move_to([5, 5, 0])
cc("disjoint_xy", a, r1)
out["collisions"] = cases
(ROOT / "tests" / "golden" / "load_zmx.json").write_text(json.dumps(out))
print("wrote load_zmx.json:", len(out["microscope"]["elements"]), "microscope elements,", out["benchmark"]["n_surfaces"], "benchmark surfaces,",
      [(c["name"], c["coll"], c["n"]) for c in cases])
