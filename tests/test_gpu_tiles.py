"""Sparse transport of detector images (otb_tiles.cu): mask / pack / unpack round trip, the tile-wise download behind
RenderImage.data, capacity overflow falling back to the dense copy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ot():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import optrace_b200 as ot
    from optrace_b200 import engine
    engine.ensure_init()
    ot.global_options.show_warnings = False
    return ot


def test_tile_pack_round_trip(ot):
    import torch
    from optrace_b200 import engine
    lib = engine.ensure_init()
    rng = np.random.default_rng(2)
    Ny, Nx = 945, 2835                                  # not multiples of the tile size: padded edge tiles
    img = np.zeros((Ny, Nx, 4))
    ys, xs = rng.integers(0, Ny, 400), rng.integers(0, Nx, 400)
    img[ys, xs] = rng.random((400, 4))
    img[Ny - 1, Nx - 1] = 1.0                           # last pixel: the corner tile
    d = torch.from_numpy(img).to(engine.device())
    tp = engine.TilePack(lib, d, 1024)
    tp.make_mask()
    tp.pack()
    h = tp.header.cpu().numpy()
    ntx = -(-Nx//engine.TILE)
    expect = np.unique((ys//engine.TILE)*ntx + xs//engine.TILE)
    expect = np.unique(np.append(expect, ((Ny - 1)//engine.TILE)*ntx + (Nx - 1)//engine.TILE))
    assert h[0] == expect.shape[0] and h[1] == 0 and np.array_equal(h[2:2 + h[0]], expect)
    host, _ = engine.assemble_tiles(d.shape, h, tp.packed.cpu().numpy())
    assert np.array_equal(host, img)
    # unpack writes the (here: doubled) tiles back
    tp.packed.mul_(2.0)
    tp.unpack()
    assert np.array_equal(d.cpu().numpy(), 2*img)
    # capacity overflow: flagged, image untouched
    tp2 = engine.TilePack(lib, d, 16)
    tp2.make_mask()
    tp2.pack()
    h2 = tp2.header[:2].cpu().numpy()
    assert h2[0] == expect.shape[0] and h2[1] == 1
    tp2.packed.zero_()
    tp2.unpack()
    assert np.array_equal(d.cpu().numpy(), 2*img)


def test_download_moves_tiles_only(ot):
    """RenderImage.data / download_async: the dense host array is rebuilt from the occupied tiles; a spread image
    (more tiles than the learnt capacity) takes the dense copy"""
    import scenes
    from optrace_b200 import engine
    RT = scenes.double_gauss(ot)
    RT.trace(400_000)
    a = RT.detector_image()
    dense = a._data_dev.cpu().numpy()
    a.download_async()
    got = a.data
    assert np.array_equal(got, dense)
    assert 0 < a.transferred_bytes < 0.3*dense.nbytes
    # second image of the same shape: capacity learnt from the first
    b = RT.detector_image()
    b.download_async()
    assert np.array_equal(b.data, b._data_dev.cpu().numpy()) and b.transferred_bytes <= a.transferred_bytes
    # a spread image: uniform illumination of the whole detector
    RT2 = ot.Raytracer(outline=[-5, 5, -5, 5, -1, 10])
    RT2.add(ot.RaySource(ot.RectangularSurface(dim=[4, 4]), pos=[0, 0, 0]))
    RT2.add(ot.Detector(ot.RectangularSurface(dim=[4, 4]), pos=[0, 0, 5]))
    RT2.trace(2_000_000)
    engine._tile_cap.pop((945, 945), None)
    c = RT2.detector_image()
    c.download_async()
    assert np.array_equal(c.data, c._data_dev.cpu().numpy()) and c.transferred_bytes >= c._data_dev.numel()*8
