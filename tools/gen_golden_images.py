"""Golden vectors for RenderImage.get, generated with the REFERENCE itself.

For the detector image of selected fixture scenes (tests/golden/<scene>.npz, variant det0) the reference's
RenderImage is rebuilt from the stored histogram and `get(mode, N, L_th, chroma_scale)` is evaluated for every
image mode; the converted arrays are stored in tests/golden/images_<scene>.npz.  Runs in the build container only
(needs /root/reference, cv2); the tests read the committed fixtures.
Usage: python tools/gen_golden_images.py"""
import pathlib
import sys
import warnings

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "tests"))
from refharness import import_reference  # noqa: E402
import golden_util as gu  # noqa: E402

CASES = {   # scene -> list of (mode key, N, L_th, chroma_scale)
    "double_gauss": 45, "image_render": 45, "arizona_eye": 63, "spherical_aberration": 45,
}


def main():
    ot = import_reference()
    warnings.simplefilter("ignore")
    for scene, N in CASES.items():
        g = gu.load(scene)
        shape = tuple(int(v) for v in g["det0_shape"])
        data = np.zeros(shape)
        data[g["det0_yi"], g["det0_xi"]] = g["det0_vals"]
        img = ot.RenderImage(extent=g["det0_extent"])
        img._data = data
        img.extent = np.array(g["det0_extent"], dtype=np.float64)
        out = {"N": N, "modes": np.array(ot.RenderImage.image_modes)}
        for k, mode in enumerate(ot.RenderImage.image_modes):
            out[f"m{k}"] = img.get(mode, N).data
        out["perc_lth"] = img.get("sRGB (Perceptual RI)", N, L_th=0.01).data
        out["perc_cs"] = img.get("sRGB (Perceptual RI)", N, chroma_scale=0.5).data
        out["abs_full"] = img.get("sRGB (Absolute RI)", 189).data       # another factor (5)
        out["irr_945"] = None
        del out["irr_945"]
        path = ROOT / "tests" / "golden" / f"images_{scene}.npz"
        np.savez_compressed(path, **out)
        inv = int(out["m2"].sum())
        print(f"{scene}: shape {shape} -> N={N}, out-of-gamut pixels {inv}, {path.name} {path.stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    main()
