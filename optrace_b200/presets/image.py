"""Procedural test images for image sources.

The reference ships photographs and charts (optrace/resources/images, third-party licences) that are NOT
redistributed here.  These generators produce synthetic stand-ins with the same pixel dimensions and call
signature (`s` side lengths or `extent`), which is what the benchmark configs need: an RGB pixel grid that
drives the pixel-CDF / sRGB-primary sampling of the device ray generator.
"""
import numpy as np

from ..images import RGBImage, GrayscaleImage


def _chart(h, w, inverted):
    """ETDRS-like chart: rows of blocky optotypes shrinking towards the bottom"""
    img = np.zeros((h, w), dtype=np.float64)
    rng = np.random.default_rng(1264)
    y = int(0.06*h)
    size = 0.11*h
    while size > 4 and y + size < h*0.97:
        n = 5
        gap = size
        x0 = (w - (n*size + (n - 1)*gap))/2
        for i in range(n):
            xs = int(x0 + i*(size + gap))
            cell = rng.random((5, 5)) > 0.45
            cell[:, 0] = True
            blk = np.kron(cell, np.ones((int(size/5) + 1, int(size/5) + 1)))[:int(size), :int(size)]
            img[y:y + blk.shape[0], xs:xs + blk.shape[1]] = blk
        y += int(size*1.9)
        size /= 1.2589
    img = np.flipud(img)
    if not inverted:
        img = 1 - img
    return np.repeat(img[:, :, None], 3, axis=2)


def ETDRS_chart(s=None, extent=None) -> RGBImage:
    return RGBImage(_chart(1200, 1264, False), s, extent, desc="ETDRS Chart (synthetic)")


def ETDRS_chart_inverted(s=None, extent=None) -> RGBImage:
    return RGBImage(_chart(1200, 1264, True), s, extent, desc="ETDRS Chart inverted (synthetic)")


def tv_testcard2(s=None, extent=None) -> RGBImage:
    """768 x 576 colour test card: colour bars, grey ramp, circle and a fine grid"""
    h, w = 576, 768
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, 3))
    bars = np.array([[1, 1, 1], [1, 1, 0], [0, 1, 1], [0, 1, 0], [1, 0, 1], [1, 0, 0], [0, 0, 1], [0.05, 0.05, 0.05]])
    img[:] = bars[np.clip(xx*8//w, 0, 7)]
    ramp = (yy > 0.72*h) & (yy < 0.86*h)
    img[ramp] = (xx[ramp]/w)[:, None]
    grid = ((xx % 48 < 2) | (yy % 48 < 2)) & (yy < 0.15*h)
    img[grid] = 1.0
    circ = np.abs(np.hypot(xx - w/2, yy - h/2) - 0.42*h) < 3
    img[circ] = 1.0
    return RGBImage(np.flipud(img), s, extent, desc="TV test card (synthetic)")


tv_testcard1 = tv_testcard2


def siemens_star(s=None, extent=None) -> RGBImage:
    h = w = 1000
    yy, xx = np.mgrid[0:h, 0:w]
    phi = np.arctan2(yy - h/2, xx - w/2)
    v = ((np.floor(phi/(2*np.pi)*72) % 2) == 0) & (np.hypot(xx - w/2, yy - h/2) < 0.48*h)
    return RGBImage(np.repeat(v[:, :, None].astype(np.float64), 3, axis=2), s, extent, desc="Siemens star")


def grid(s=None, extent=None) -> RGBImage:
    h = w = 801
    yy, xx = np.mgrid[0:h, 0:w]
    v = ((xx % 80 < 3) | (yy % 80 < 3)).astype(np.float64)
    return RGBImage(np.repeat(v[:, :, None], 3, axis=2), s, extent, desc="Grid")


def color_checker(s=None, extent=None) -> RGBImage:
    rng = np.random.default_rng(24)
    patches = rng.random((4, 6, 3))
    return RGBImage(np.kron(patches, np.ones((100, 100, 1))), s, extent, desc="Colour patches (synthetic)")


all_presets = [ETDRS_chart, ETDRS_chart_inverted, tv_testcard1, tv_testcard2, siemens_star, grid, color_checker]
