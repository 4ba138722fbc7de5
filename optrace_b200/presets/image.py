"""Image presets (optrace/tracer/presets/image.py): the test charts and photographs the reference ships for image
sources, loaded from optrace_b200/data/images/.  The files are the reference's own resource images (public
domain / free-to-use material, sources and licences in data/images/SOURCE.txt) — data like the glass catalogue and
the CIE tables, not code; each function keeps the reference's name, signature and description string.

`s` = side lengths [sx, sy] in mm or `extent` = [x0, x1, y0, y1], as for RGBImage / GrayscaleImage."""
import pathlib

import numpy as np

from ..images import RGBImage, GrayscaleImage

IMAGE_DIR = pathlib.Path(__file__).resolve().parent.parent / "data" / "images"


def _rgb(file: str, desc: str):
    def preset(s=None, extent=None) -> RGBImage:
        return RGBImage(str(IMAGE_DIR / file), s, extent, desc=desc)
    preset.__name__ = preset.__qualname__ = file.split(".")[0]
    preset.__doc__ = f"{desc} (data/images/{file}; presets/image.py of the reference)"
    return preset


def _gray(file: str, desc: str):
    def preset(s=None, extent=None) -> GrayscaleImage:
        return RGBImage(str(IMAGE_DIR / file), s, extent, desc=desc).to_grayscale_image()
    preset.__name__ = preset.__qualname__ = file.split(".")[0]
    preset.__doc__ = f"{desc} as a GrayscaleImage (data/images/{file})"
    return preset


# photographs (presets/image.py:13-88)
cell = _rgb("cell.webp", "Cell")
documents = _rgb("documents.webp", "Documents")
fruits = _rgb("fruits.webp", "Fruits")
group_photo = _rgb("group_photo.webp", "Group Photo")
hong_kong = _rgb("hong_kong.webp", "Hong Kong")
interior = _rgb("interior.webp", "Interior")
landscape = _rgb("landscape.webp", "Landscape")
scenes = [cell, documents, fruits, group_photo, hong_kong, interior, landscape]

# test charts (presets/image.py:97-191)
color_checker = _rgb("color_checker.webp", "Color Checker Chart")
ETDRS_chart = _gray("ETDRS_chart.png", "ETDRS Chart")
ETDRS_chart_inverted = _gray("ETDRS_chart_inverted.png", "ETDRS Chart Inverted")
eye_test_vintage = _rgb("eye_test_vintage.webp", "Eye Test Vintage")
siemens_star = _gray("siemens_star.png", "Siemens Star")
tv_testcard1 = _rgb("tv_testcard1.png", "TV Testcard 1")
tv_testcard2 = _rgb("tv_testcard2.png", "TV Testcard 2")


def grid(s=None, extent=None) -> GrayscaleImage:
    """white grid of 10 x 10 cells on black, 301 x 301 px (presets/image.py:142-155)"""
    g = np.zeros((301, 301))
    g[::30] = 1
    g[:, ::30] = 1
    return GrayscaleImage(g, s, extent, desc="Grid")


test_images = [color_checker, ETDRS_chart, ETDRS_chart_inverted, eye_test_vintage, grid, siemens_star, tv_testcard1,
               tv_testcard2]
all_presets = [*test_images, *scenes]
