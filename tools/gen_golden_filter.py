"""Golden vectors for the resolution-limit filter, generated with the REFERENCE itself:
RenderImage.render(p, w, wl, limit=...) on the detector hits of a fixture scene.  Stored: the filtered image
irradiance joined to 189 x 189 bins, a 96 x 96 full-resolution crop around the brightest pixel, the
enlarged extent and the channel sums.  Build container only (needs /root/reference)."""
import pathlib
import sys
import warnings

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "tests"))
from refharness import import_reference  # noqa: E402
import golden_util as gu  # noqa: E402

CASES = {"spherical_aberration": 60.0, "image_render": 25.0}      # limit in micrometres


def main():
    ot = import_reference()
    ot.global_options.multithreading = False
    warnings.simplefilter("ignore")
    for scene, limit in CASES.items():
        g = gu.load(scene)
        ph, w, wl = g["det0_ph"], g["det0_w"], g["det0_wl"]
        img = ot.RenderImage(extent=g["det0_extent0"])
        img.render(ph, w, wl, limit=limit)
        d = img._data
        Ny, Nx, _ = d.shape
        y0, x0 = np.unravel_index(np.argmax(d[:, :, 3]), d[:, :, 3].shape)
        y0, x0 = int(np.clip(y0 - 48, 0, Ny - 96)), int(np.clip(x0 - 48, 0, Nx - 96))
        out = dict(limit=limit, extent=img.extent, shape=np.array(d.shape), sums=d.sum(axis=(0, 1)),
                   crop_origin=np.array([y0, x0]), crop=d[y0:y0 + 96, x0:x0 + 96].copy(),
                   irr=img.get("Irradiance", 189).data)
        path = ROOT / "tests" / "golden" / f"filter_{scene}.npz"
        np.savez_compressed(path, **out)
        print(scene, d.shape, "extent", img.extent, "->", path.name, f"{path.stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    main()
