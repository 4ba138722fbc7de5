"""ZEMAX importers of the drop-in API (optrace/tracer/load.py): `.agf` glass catalogues -> RefractionIndex
dictionaries (load.py:57-144) and sequential `.zmx` lens files -> Group (load.py:147-416).  Host-side file
parsing at scene-construction time: it produces the same scene objects the reference
builds, so that the reference's own benchmark scene (tests/benchmark.py: microscope + eyepiece + eye, 57 surfaces)
runs on this engine.

Both formats are line oriented: a two / four letter record tag followed by whitespace separated fields.  The
parsers below dispatch on the tag through small tables; supported records, defaults and error messages follow the
reference (cited per function)."""
from __future__ import annotations

import os

import numpy as np

from .elements import Aperture, Detector, Group, Lens, PointMarker
from .media import RefractionIndex
from .options import warning
from .presets import spectral_lines
from .surfaces import AsphericSurface, CircularSurface, ConicSurface, RectangularSurface, RingSurface, SphericalSurface, Surface

# formula number in an .agf "NM" record -> RefractionIndex n_type (load.py:19-21; ZEMAX manual, glass catalogs)
AGF_FORMULAS = ("Schott", "Sellmeier1", "Herzberger", "Sellmeier2", "Conrady", "Sellmeier3", "Handbook of Optics 1",
                "Handbook of Optics 2", "Sellmeier4", "Extended", "Sellmeier5", "Extended2", "Extended3")


def _text_lines(path: str) -> list[str]:
    """file -> list of lines (load.py:24-54).  The encoding is sniffed from the byte-order mark: ZEMAX writes
    UTF-16 with BOM or plain 8-bit text (the reference asks chardet, which gives the same answer for such files)."""
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found/ is not a file.")
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] in (b"\xff\xfe", b"\xfe\xff"):
        text = raw.decode("utf-16")
    elif raw[:3] == b"\xef\xbb\xbf":
        text = raw[3:].decode("utf-8")
    else:
        try:
            text = raw.decode("utf-8")
        except UnicodeDecodeError:
            text = raw.decode("latin-1")
    # universal newlines like a text-mode read: every record keeps exactly one trailing "\n"
    text = text.lstrip("﻿").replace("\r\n", "\n").replace("\r", "\n")
    return text.splitlines(keepends=True)


def _index_check(name: str, n: RefractionIndex, n_file: float, V_file: float, lo_nm: float, hi_nm: float) -> None:
    """load.py:116-136: index at the d line and Abbe number against the catalogue's own values (warnings only).
    n(lambda) is evaluated by the engine; without a CUDA device the comparison is skipped, the material is kept."""
    F, d, C = spectral_lines.FdC
    if lo_nm > F or hi_nm < C:
        warning(f"{name} wavelength range [{lo_nm}, {hi_nm}]nm does not overlap with "
                f"testing wavelengths {spectral_lines.FdC}nm, skipping index and Abbe number checks.")
        return
    from ._cabi import EngineError
    try:
        n_d = float(np.atleast_1d(n(spectral_lines.d))[0])
        V = n.abbe_number(spectral_lines.FdC)
    except EngineError:
        return
    if abs(n_d - n_file) > 1e-4:
        warning(f"{name}: Index from file is {n_file}, but calculated index is {n_d}. "
                "This can be due to different probe wavelengths.")
    elif abs(V - V_file) > 0.3:
        warning(f"{name}: The Abbe number from file is {V_file}, but calculated is {V}. "
                "This can be due to different probe wavelengths.")


def load_agf(path: str) -> dict:
    """load.py:57-144: `.agf` catalogue -> {material name: RefractionIndex}.
    Records used: NM <name> <formula> <..> <nd> <Vd> ...;  CD <coefficients>;  LD <min um> <max um>.
    A material is complete at its LD record; unknown formula numbers and materials whose index model is invalid
    (e.g. n < 1 somewhere) are skipped with a warning."""
    out: dict = {}
    cur = None            # material under construction: dict(name, mode, nd, Vd, coeff)
    for line in _text_lines(path):
        tag, fields = line[:2], line.split()[1:]
        if tag == "NM":
            name = fields[0]
            num = int(float(fields[1]))
            if not 1 <= num <= len(AGF_FORMULAS):
                warning(f"{name}: Unknown index formula mode number {num}, skipping.")
                cur = None
                continue
            cur = dict(name=name, mode=AGF_FORMULAS[num - 1], nd=float(fields[3]), Vd=float(fields[4]), coeff=None)
        elif cur is None:
            continue
        elif tag == "CD":
            need = RefractionIndex.coeff_count[cur["mode"]]
            c = [float(v) for v in fields][:need]
            cur["coeff"] = c + [0.0]*(need - len(c))
        elif tag == "LD":
            try:
                n = RefractionIndex(cur["mode"], coeff=cur["coeff"], desc=cur["name"])
                _index_check(cur["name"], n, cur["nd"], cur["Vd"], float(fields[0])*1000, float(fields[1])*1000)
                out[cur["name"]] = n
            except Exception as err:        # noqa: BLE001 — the reference skips any material that fails (load.py:141-142)
                warning(f"Error for material {cur['name']}: " + str(err))
    return out


# ---- .zmx -------------------------------------------------------------------------------------------------------
def _glass(fields: list[str], n_dict: dict) -> RefractionIndex:
    """GLAS record (load.py:282-305): catalogue material, or an Abbe model from the (nd, Vd) the record carries"""
    name = fields[0]
    nd, Vd = (float(fields[3]), float(fields[4])) if len(fields) > 5 else (None, None)
    if name == "___BLANK":
        return RefractionIndex("Abbe", n=nd, V=Vd)
    if name in n_dict:
        return n_dict[name]
    if nd is not None and Vd is not None and nd > 1 and Vd > 0:
        return RefractionIndex("Abbe", n=nd, V=Vd)
    raise RuntimeError(f"Material {name} missing in n_dict parameter.")


def _parse_zmx(lines: list[str], n_dict: dict):
    """load.py:197-326: header checks, then one property dict per SURF block.
    Returns (surface dicts, thickness behind each surface, ambient medium or None, file description)."""
    desc, k = "", 0
    for k, line in enumerate(lines):
        tag = line[:4]
        if tag == "NAME":
            desc = line[5:-1]
        elif tag == "UNIT" and line.split()[1] != "MM":
            raise RuntimeError(f"Unsupported Unit {line.split()[1]}.")
        elif tag == "MODE" and line.split()[1] != "SEQ":
            raise RuntimeError(f"Unsupported Mode {line.split()[1]}.")
        elif tag == "SURF":
            break
    # blocks: the lines between consecutive SURF records.  Like the reference's scan (load.py:229-233) the very last
    # line of the file is never read as a record, and a block is only closed by a following line
    blocks, i = [], k + 1
    while i < len(lines):
        block = []
        while i + 1 < len(lines) and lines[i][:4] != "SURF":
            block.append(lines[i])
            i += 1
        blocks.append(block)
        i += 1

    surfaces, gaps, n0 = [], [], None
    for num, block in enumerate(blocks):
        s = dict(stype="STANDARD", desc="", k=0, R=np.inf)
        parm, gap = [0.0]*10, 0.0
        for line in block:
            tag, f = line[2:6], line.split()[1:]
            if tag == "TYPE":
                s["stype"] = f[0]
            elif tag == "DIAM":
                s["r"] = max(float(f[0]), 1e-9)
            elif tag == "CONI":
                s["k"] = float(f[0])
            elif tag == "COMM":
                s["desc"] = line[7:-1]
            elif tag == "COAT":
                warning(f"Coatings are not supported. Ignoring coating '{line[7:-1]}'.")
            elif tag == "STOP":
                s["STOP"] = True
            elif tag == "CURV":
                rho = float(f[0])
                s["R"] = 1/rho if rho else np.inf
            elif tag == "DISZ":
                gap = max(float(f[0]), 3*Surface.N_EPS)        # touching surfaces keep a minimal distance
            elif tag == "PARM":
                parm[int(float(f[0])) - 1] = float(f[1])
            elif tag == "GLAS":
                s["n"] = _glass(f, n_dict)
        if num == 0 and not np.isfinite(gap):
            # object surface at infinity: only its medium matters (ambient index in front of the system)
            n0 = s.get("n", RefractionIndex("Constant", n=1))
        else:
            s["parm"] = parm
            surfaces.append(s)
            gaps.append(gap)
    return surfaces, gaps, n0, desc


def _surface_of(s: dict) -> Surface:
    """load.py:172-194: STANDARD -> circle / sphere / conic, EVENASPH -> asphere"""
    if s["stype"] == "STANDARD":
        if not np.isfinite(s["R"]):
            return CircularSurface(r=s["r"], desc=s["desc"])
        if s.get("k"):
            return ConicSurface(r=s["r"], R=s["R"], k=s["k"], desc=s["desc"])
        return SphericalSurface(r=s["r"], R=s["R"], desc=s["desc"])
    if s["stype"] == "EVENASPH":
        return AsphericSurface(r=s["r"], R=s["R"], k=s["k"], coeff=s["parm"], desc=s["desc"])
    raise RuntimeError("Surface mode " + str(s["stype"]) + " not supported yet.")


def _assemble(surfaces: list, gaps: list, n0, desc: str, no_marker: bool) -> Group:
    """load.py:328-416: a surface that carries a glass opens a lens that the next surface closes; when that next
    surface carries a glass too, it also opens the following lens (cemented group: the second lens starts 1e-7 mm
    behind, the gap keeps the first lens' medium).  Glass-free surfaces are the stop (ring aperture), the image
    plane (last surface -> square detector) or plain spacing."""
    G = Group(long_desc=desc, n0=n0)
    r_max = max([s["r"] for s in surfaces if "r" in s], default=0)
    for s in surfaces:
        s.setdefault("r", r_max)              # laterally unbounded media take the largest radius of the file
    i = next((j for j, s in enumerate(surfaces) if "n" in s), len(surfaces))
    z = 0.0
    while i < len(surfaces):
        s = surfaces[i]
        if "n" not in s:
            if i + 1 == len(surfaces) and "r" in s:
                G.add(Detector(RectangularSurface(dim=[2*s["r"], 2*s["r"]]), pos=[0, 0, z], desc=s["desc"]))
            elif "STOP" in s:
                e = G.extent
                r = max(s["r"] + 1, max(e[1] - e[0], e[3] - e[2])/2)
                G.add(Aperture(RingSurface(ri=s["r"], r=r), pos=[0, 0, z], desc=s["desc"]))
            z += gaps[i]
            i += 1
            continue
        nxt = surfaces[i + 1]
        cemented = "n" in nxt
        G.add(Lens(_surface_of(s), _surface_of(nxt), n=s["n"], pos=[0, 0, z], d1=0, d2=gaps[i],
                   n2=s["n"] if cemented else RefractionIndex("Constant", n=1), desc=s["desc"]))
        if cemented:
            z += gaps[i] + 1e-7
            i += 1
        else:
            z += gaps[i] + gaps[i + 1]
            i += 2
    if G.long_desc != "" and not no_marker:
        e = G.extent
        G.add(PointMarker(G.long_desc, [e[0] - 1.5, np.mean(e[2:4]), np.mean(e[4:6])], label_only=True))
    return G


def load_zmx(filename: str, n_dict: dict = None, no_marker: bool = False) -> Group:
    """load.py:147-169: sequential `.zmx` file (units mm) -> Group of lenses, stop aperture and image detector"""
    surfaces, gaps, n0, desc = _parse_zmx(_text_lines(filename), n_dict or {})
    return _assemble(surfaces, gaps, n0, desc, no_marker)
