"""Build optrace_b200/data/cie_tables.npz from the CIE data tables shipped with the reference.

The tables are published CIE standard data (CIE 1931 2-degree colour matching functions,
DOI 10.25039/CIE.DS.xvudnb9b; CIE illuminants, "CIE Colorimetry, 3rd edition, 2004" and
DOI 10.25039/CIE.DS.vgssnyfg) — numerical standards, not reference source code.  Parsed exactly like
optrace/tracer/color/observers.py:11-12 and illuminants.py:9-13 (np.genfromtxt, empty cells -> 0) so the
float64 values are bit-identical to what the reference interpolates.

Run in the development container only (needs /root/reference):  python tools/extract_cie_tables.py
"""
import pathlib
import numpy as np

REF = pathlib.Path("/root/reference/optrace/resources")
OUT = pathlib.Path(__file__).resolve().parent.parent / "optrace_b200" / "data" / "cie_tables.npz"

obs = np.genfromtxt(REF / "observers.csv", skip_header=1, delimiter=",", filling_values=0, dtype=np.float64)
ill = np.genfromtxt(REF / "illuminants.csv", skip_header=1, delimiter=",", filling_values=0, dtype=np.float64)
ill_names = ["wl", "A", "C", "D50", "D55", "D65", "D75", "F2", "F7", "F11", "LED-B1", "LED-B2", "LED-B3",
             "LED-B4", "LED-B5", "LED-BH1", "LED-RGB1", "LED-V1", "LED-V2"]
np.savez_compressed(OUT, observers=obs, illuminants=ill, illuminant_names=np.array(ill_names))
print("wrote", OUT, obs.shape, ill.shape)
