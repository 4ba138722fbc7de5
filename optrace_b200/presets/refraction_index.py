"""Preset materials built from the catalogue table data/glass_catalog.json (manufacturer dispersion data;
reference: presets/refraction_index.py)."""
import json as _json
import pathlib as _pathlib

import numpy as _np

from ..media import RefractionIndex

_cat = _json.loads((_pathlib.Path(__file__).resolve().parent.parent / "data" / "glass_catalog.json").read_text())


def _make(e):
    kw = dict(desc=e["desc"], long_desc=e["long_desc"])
    t = e["n_type"]
    if t == "Constant":
        return RefractionIndex("Constant", n=e["n"], **kw)
    if t == "Abbe":
        return RefractionIndex("Abbe", n=e["n"], V=e["V"], **kw)
    if t == "Data":
        return RefractionIndex("Data", wls=_np.array(e["wls"]), vals=_np.array(e["vals"]), **kw)
    return RefractionIndex(t, coeff=list(e["coeff"]), **kw)


for _name, _e in _cat["materials"].items():
    globals()[_name] = _make(_e)

glasses = [globals()[n] for n in _cat["groups"]["glasses"]]
plastics = [globals()[n] for n in _cat["groups"]["plastics"]]
misc = [globals()[n] for n in _cat["groups"]["misc"]]
all_presets = [*glasses, *plastics, *misc]
