"""Build optrace_b200/data/glass_catalog.json: dispersion-formula coefficients of the preset materials.

The numbers are manufacturer / literature catalogue data (SCHOTT data sheets, refractiveindex.info, as cited
per entry in the reference's presets/refraction_index.py) — physical constants, not program code.  They are
read through the importable reference so that the drop-in presets evaluate to bit-identical indices.
Materials defined by a Python callable in the reference ("Function" type) are skipped.

Run in the development container only:  python tools/extract_presets.py
"""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
from refharness import import_reference

ot = import_reference()
mod = ot.presets.refraction_index
out = {}
for name in dir(mod):
    obj = getattr(mod, name)
    if not isinstance(obj, ot.RefractionIndex) or obj.spectrum_type == "Function":
        continue
    e = dict(n_type=obj.spectrum_type, desc=obj.desc, long_desc=obj.long_desc)
    if obj.spectrum_type == "Constant":
        e["n"] = obj.val
    elif obj.spectrum_type == "Abbe":
        e["n"], e["V"] = obj.val, obj.V
    elif obj.spectrum_type == "Data":
        e["wls"], e["vals"] = [float(v) for v in obj._wls], [float(v) for v in obj._vals]
    else:
        e["coeff"] = [float(v) for v in obj.coeff]
    out[name] = e
groups = {g: [n for n in out if any(getattr(mod, n) is o for o in getattr(mod, g))]
          for g in ("glasses", "plastics", "misc")}
path = pathlib.Path(__file__).resolve().parent.parent / "optrace_b200" / "data" / "glass_catalog.json"
path.write_text(json.dumps(dict(materials=out, groups=groups), indent=1))
print("wrote", path, len(out), "materials")
