"""ZEMAX importers (optrace/tracer/load.py:57-416) and Raytracer.check_collision / geometry checks
(raytracer.py:510-664) against what the reference itself builds from its benchmark files and answers on surface
pairs (tests/golden/load_zmx.json, generator tools/gen_golden_load.py).  The .agf / .zmx resource files travel with
the vendored reference (oracle/_ref, tools/vendor_reference.py); without them the file-based tests skip."""
import json
import pathlib

import numpy as np
import pytest

import optrace_b200 as ot
import golden_util as gu

RES = gu.ROOT / "oracle" / "_ref" / "examples" / "resources"
G = json.loads((gu.GOLDEN / "load_zmx.json").read_text())
needs_files = pytest.mark.skipif(not RES.exists(), reason="reference resource files not vendored (oracle/_ref)")


def _medium_eq(n, d):
    if d is None:
        return n is None
    if n.spectrum_type != d["type"]:
        return False
    if d["coeff"] is not None and [float(v) for v in n.coeff] != d["coeff"]:
        return False
    if d["type"] in ("Constant", "Abbe") and float(n.val) != d["val"]:
        return False
    return d["V"] is None or float(n.V) == d["V"]


def _surface_eq(s, d):
    assert type(s).__name__ == d["cls"], (type(s).__name__, d["cls"])
    assert float(s.r) == d["r"] and np.allclose(s.pos, d["pos"], rtol=0, atol=1e-12)
    assert abs(s.z_min - d["z_min"]) < 1e-12 and abs(s.z_max - d["z_max"]) < 1e-12
    for k in ("R", "k", "ri"):
        if k in d:
            assert float(getattr(s, k)) == d[k], k
    if "coeff" in d:
        assert [float(v) for v in s.coeff] == d["coeff"]
    if "dim" in d:
        assert [float(v) for v in s.dim] == d["dim"]


def _group_eq(Gr, d):
    assert _medium_eq(Gr.n0, d["n0"]) and Gr.long_desc == d["long_desc"]
    els = Gr.elements
    assert len(els) == len(d["elements"])
    for el, e in zip(els, d["elements"]):
        assert type(el).__name__ == e["cls"] and el.desc == e["desc"]
        assert np.allclose(el.pos, e["pos"], rtol=0, atol=1e-12)
        _surface_eq(el.front, e["front"])
        if "back" in e:
            _surface_eq(el.back, e["back"])
            assert abs(el.d1 - e["d1"]) < 1e-12 and abs(el.d2 - e["d2"]) < 1e-12
        if "n" in e:
            assert _medium_eq(el.n, e["n"]) and _medium_eq(el.n2, e["n2"])


@pytest.fixture(scope="module")
def n_dict():
    ot.global_options.show_warnings = False
    d = {}
    for f in ("schott", "ohara", "hikari", "hoya"):
        d |= ot.load_agf(str(RES / "materials" / f"{f}.agf"))
    return d


@needs_files
def test_load_agf_matches_reference(n_dict):
    for f in ("schott", "ohara", "hikari", "hoya"):
        ours = ot.load_agf(str(RES / "materials" / f"{f}.agf"))
        ref = G["catalogues"][f]
        assert set(ours) == set(ref), f
        for k, v in ref.items():
            assert _medium_eq(ours[k], v), (f, k)
    with pytest.raises(FileNotFoundError):
        ot.load_agf("no_such_file.agf")


@needs_files
def test_load_zmx_matches_reference(n_dict):
    _group_eq(ot.load_zmx(str(RES / "microscope" / "Nikon_1p25NA_60x_US7889433B2_MultiConfig_v2.zmx"), n_dict=n_dict), G["microscope"])
    _group_eq(ot.load_zmx(str(RES / "eyepiece" / "UK565851-1.zmx"), n_dict=n_dict), G["eyepiece"])
    # materials missing in the dictionary fall back to the Abbe model of the (nd, Vd) the file carries
    E = ot.load_zmx(str(RES / "eyepiece" / "UK565851-1.zmx"), n_dict={})
    assert all(L.n.spectrum_type == "Abbe" for L in E.lenses)


def test_zmx_parser_records(tmp_path):
    """hand-written file: unit / mode checks, conic + asphere + stop + image surfaces, cemented pair, blank glass"""
    txt = "\n".join([
        "VERS 1", "MODE SEQ", "NAME test system", "UNIT MM X W X CM MR CPMM",
        "SURF 0", "  TYPE STANDARD", "  CURV 0.0", "  DISZ INFINITY",
        "SURF 1", "  TYPE STANDARD", "  CURV 0.05 0 0 0 0", "  DISZ 2.0", "  GLAS ___BLANK 1 0 1.5 60.0 0 0 0 0 0 0", "  DIAM 5.0 0 0 0 1", "  COMM front",
        "SURF 2", "  TYPE STANDARD", "  CURV -0.04 0 0 0 0", "  CONI -0.5", "  DISZ 1.0", "  GLAS ___BLANK 1 0 1.7 30.0 0 0 0 0 0 0", "  DIAM 5.0 0 0 0 1",
        "SURF 3", "  TYPE EVENASPH", "  CURV 0.01 0 0 0 0", "  PARM 1 0.0", "  PARM 2 1e-4", "  DISZ 3.0", "  DIAM 4.5 0 0 0 1",
        "SURF 4", "  STOP", "  TYPE STANDARD", "  CURV 0.0", "  DISZ 10.0", "  DIAM 2.0 0 0 0 1",
        "SURF 5", "  TYPE STANDARD", "  CURV 0.0", "  DISZ 0.0", "  DIAM 6.0 0 0 0 1", "  COMM image", "END", ""])
    f = tmp_path / "t.zmx"
    f.write_text(txt)
    Gr = ot.load_zmx(str(f))
    assert Gr.long_desc == "test system" and len(Gr.lenses) == 2 and len(Gr.apertures) == 1 and len(Gr.detectors) == 1
    L0, L1 = Gr.lenses
    assert type(L0.front).__name__ == "SphericalSurface" and L0.front.R == 20.0 and type(L0.back).__name__ == "ConicSurface"
    assert L0.back.k == -0.5 and L0.n2 is L0.n and L1.n.val == 1.7            # cemented: gap keeps the first medium
    assert type(L1.front).__name__ == "ConicSurface" and type(L1.back).__name__ == "AsphericSurface"
    assert abs(L1.pos[2] - (2.0 + 1e-7)) < 1e-12 and L1.back.coeff[1] == 1e-4
    assert Gr.apertures[0].surface.ri == 2.0 and abs(Gr.apertures[0].pos[2] - (2.0 + 1e-7 + 1.0 + 3.0)) < 1e-12
    assert Gr.detectors[0].desc == "image" and [float(v) for v in Gr.detectors[0].surface.dim] == [12.0, 12.0]
    f.write_text(txt.replace("UNIT MM", "UNIT IN"))
    with pytest.raises(RuntimeError):
        ot.load_zmx(str(f))
    f.write_text(txt.replace("MODE SEQ", "MODE NSC"))
    with pytest.raises(RuntimeError):
        ot.load_zmx(str(f))
    # a material that is neither in the dictionary nor described by (nd, Vd) in the file
    f.write_text(txt.replace("GLAS ___BLANK 1 0 1.7 30.0 0 0 0 0 0 0", "GLAS NOSUCHGLASS 0 0"))
    with pytest.raises(RuntimeError):
        ot.load_zmx(str(f))


def test_check_collision_known_answers():
    S = ot.SphericalSurface
    a = S(r=3, R=5); a.move_to([0, 0, 0])
    b = S(r=3, R=-5); b.move_to([0, 0, 1.2])
    b2 = S(r=3, R=-5); b2.move_to([0, 0, 2.5])
    c1 = ot.CircularSurface(r=2); c1.move_to([0.5, 0, 0.3])
    t = ot.TiltedSurface(r=3, normal=[0.4, 0, 1]); t.move_to([0, 0, 0.5])
    p = ot.Point(); p.move_to([1.0, 0.5, 0.05])
    ln = ot.Line(r=2.5, angle=30); ln.move_to([0, 0, 0.4])
    r1 = ot.RectangularSurface(dim=[2, 2]); r1.move_to([5, 5, 0])
    pairs = dict(biconvex_thin=(a, b), biconvex_ok=(a, b2), sphere_vs_circle=(a, c1), sphere_vs_tilted=(a, t),
                 point_front=(p, a), surface_point=(a, p), line_front=(ln, a), disjoint_xy=(a, r1))
    for c in G["collisions"]:
        coll, x, y, z = ot.Raytracer.check_collision(*pairs[c["name"]])
        assert coll == c["coll"] and x.shape[0] == c["n"], c["name"]
        if c["first"] is not None:
            assert np.allclose([x[0], y[0], z[0]], c["first"], rtol=0, atol=1e-12), c["name"]
    with pytest.raises(TypeError):
        ot.Raytracer.check_collision(p, ln)


def test_geometry_checks_flag_collisions_and_order():
    """raytracer.py:510-578: a lens whose surfaces intersect, elements reaching into each other and a source inside
    a lens set geometry_error (trace aborts with a warning); the fault positions are reported"""
    ot.global_options.show_warnings = False
    n = ot.RefractionIndex("Constant", n=1.5)
    RT = ot.Raytracer(outline=[-10, 10, -10, 10, -10, 30])
    RT.add(ot.RaySource(ot.CircularSurface(r=1), pos=[0, 0, -5]))
    RT.add(ot.Lens(ot.SphericalSurface(r=3, R=5), ot.SphericalSurface(r=3, R=-5), n=n, pos=[0, 0, 0], d=3.0))
    RT._geometry_checks()
    assert not RT.geometry_error
    # second lens pushed into the first one
    L2 = ot.Lens(ot.SphericalSurface(r=3, R=-4), ot.CircularSurface(r=3), n=n, pos=[0, 0, 1.6], d1=0.2, d2=0.5)
    RT.add(L2)
    RT._geometry_checks()
    assert RT.geometry_error and RT.fault_pos.shape[0] > 0 and RT.fault_pos.shape[1] == 3
    RT.remove(L2)
    RT._geometry_checks()
    assert not RT.geometry_error
    # source sitting inside the lens
    RT.ray_sources[0].move_to([0, 0, 0.0])
    RT._geometry_checks()
    assert RT.geometry_error
    # element outside the outline
    RT.ray_sources[0].move_to([0, 0, -5])
    RT.add(ot.Filter(ot.CircularSurface(r=2), pos=[0, 0, 40], spectrum=ot.TransmissionSpectrum("Constant", val=0.5)))
    RT._geometry_checks()
    assert RT.geometry_error
