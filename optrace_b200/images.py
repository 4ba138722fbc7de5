"""Image containers of the drop-in API: RGBImage / GrayscaleImage (pixel data feeding image sources) and
RenderImage (the XYZW detector histogram, rendered on the device).

References: optrace/tracer/image/base_image.py, rgb_image.py, grayscale_image.py, render_image.py.
Post-processing of finished images (sRGB conversion, rescaling, Airy filter, file export) is out of scope
 — RenderImage here holds the raw (Ny, Nx, 4) float64 histogram exactly as the
reference's `RenderImage._data`.
"""
from __future__ import annotations

import copy as _copy

import numpy as np


class _BaseImage:
    _channels = None

    def __init__(self, data, s=None, extent=None, projection: str = None, quantity: str = "",
                 limit: float = None, desc: str = "", long_desc: str = ""):
        if isinstance(data, str):
            data = self._load_image(data)
        if not isinstance(data, np.ndarray):
            raise TypeError("data needs to be a numpy array or a file path.")
        self._data = self._check_data(np.asarray_chkfinite(data, dtype=np.float64))
        if extent is None and s is None:
            raise ValueError("Either s or extent need to be provided for Images")
        if extent is None:
            s2 = np.asarray_chkfinite(s, dtype=np.float64)
            if s2.shape[0] != 2:
                raise ValueError("s needs to have 2 elements.")
            if s2[0] <= 0 or s2[1] <= 0:
                raise ValueError("s needs to be positive.")
            extent = [-s2[0]/2, s2[0]/2, -s2[1]/2, s2[1]/2]
        e = np.asarray_chkfinite(extent, dtype=np.float64)
        if e.shape[0] != 4:
            raise ValueError("Extent needs to have 4 elements.")
        if e[0] > e[1] or e[2] > e[3]:
            raise ValueError("Extent needs to be an array with [x0, x1, y0, y1] with x0 < x1 and y0 < y1.")
        self.extent = e
        self.quantity, self.projection, self.limit = quantity, projection, limit
        self.desc, self.long_desc = desc, long_desc

    def _load_image(self, path: str) -> np.ndarray:
        """base_image.py:68-86: file loading through OpenCV (host I/O, not on the ray path)."""
        import cv2
        if not cv2.haveImageReader(path):
            raise IOError(f"Can't find/process file {path}")
        image = np.flipud(cv2.imread(path, flags=cv2.IMREAD_COLOR))
        if self._channels == 3:
            return cv2.cvtColor(image, cv2.COLOR_BGR2RGB)/255.0
        return cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)/255.0

    def _check_data(self, d):
        return d

    def copy(self):
        return _copy.deepcopy(self)

    @property
    def shape(self):
        return self._data.shape

    @property
    def data(self) -> np.ndarray:
        return self._data.copy()

    @property
    def s(self):
        return [float(self.extent[1] - self.extent[0]), float(self.extent[3] - self.extent[2])]

    @property
    def Apx(self) -> float:
        return float(self.s[0]*self.s[1]/(self.shape[1]*self.shape[0]))

    def profile(self, x: float = None, y: float = None):
        """base_image.py:149-186: image cut at x or y (nearest pixel): (bin edges, list of cuts)"""
        img = self._data
        if x is not None:
            if not self.extent[0] <= x <= self.extent[1]:
                raise ValueError(f"Position x={x} is outside the image x-extent of {self.extent[:2]}")
            bins = np.linspace(self.extent[2], self.extent[3], self.shape[0] + 1)
            ind = int((x - self.extent[0])/self.s[0]*self.shape[1]*(1 - 1e-12))
            iml = [img[:, ind]] if img.ndim == 2 else [img[:, ind, 0], img[:, ind, 1], img[:, ind, 2]]
        elif y is not None:
            if not self.extent[2] <= y <= self.extent[3]:
                raise ValueError(f"Position y={y} is outside the image y-extent of {self.extent[2:]}")
            bins = np.linspace(self.extent[0], self.extent[1], self.shape[1] + 1)
            ind = int((y - self.extent[2])/self.s[1]*self.shape[0]*(1 - 1e-12))
            iml = [img[ind]] if img.ndim == 2 else [img[ind, :, 0], img[ind, :, 1], img[ind, :, 2]]
        else:
            raise ValueError("Either x or y parameter must be provided.")
        return bins, iml


class RGBImage(_BaseImage):
    """image/rgb_image.py: [0, 0] is the lower-left pixel, values in [0, 1], 3 channels."""
    _channels = 3

    def _check_data(self, d):
        if d.ndim != 3 or d.shape[2] != 3:
            raise ValueError(f"Image needs to have three dimensions with 3 elements (RGB) in the third "
                             f"dimension, but has shape {d.shape}.")
        if np.min(d) < 0.0 or np.max(d) > 1.0:
            raise ValueError("Make sure all image data is in the range [0, 1].")
        return d


    def to_grayscale_image(self) -> "GrayscaleImage":
        """rgb_image.py:41-56: luminance Y of the linearised sRGB values (Y row of the sRGB -> XYZ matrix,
        color/srgb.py:61-63), gamma-compressed back and clipped to [0, 1]"""
        from . import color
        lin = color.srgb_to_srgb_linear(self._data)
        Y = 0.2126729*lin[:, :, 0] + 0.7151522*lin[:, :, 1] + 0.0721750*lin[:, :, 2]
        low = np.abs(Y) <= 0.0031308
        g = np.where(low, 12.92*Y, np.sign(Y)*(1.055*np.abs(Y)**(1/2.4) - 0.055))
        return GrayscaleImage(np.clip(g, 0, 1), extent=self.extent, desc=self.desc, long_desc=self.long_desc,
                              quantity=self.quantity, projection=self.projection, limit=self.limit)


class ScalarImage(_BaseImage):
    """image/scalar_image.py: one channel, non-negative values (irradiance, illuminance, CIELUV channels ...)"""
    _channels = 1

    def _check_data(self, d):
        if d.ndim == 3:
            raise ValueError("Image can't have color information. Either use a RGBImage or remove color information.")
        if d.ndim != 2:
            raise ValueError(f"Image needs to have two dimensions but has shape {d.shape}.")
        if d.size and (m := d.min()) < 0.0:
            raise ValueError(f"There is a negative value of {m} inside the image. Make sure all image data is "
                             "non-negative.")
        return d


class GrayscaleImage(_BaseImage):
    """image/grayscale_image.py"""
    _channels = 1

    def _check_data(self, d):
        if d.ndim != 2:
            raise ValueError(f"Image needs to have two dimensions, but has shape {d.shape}.")
        if np.min(d) < 0.0 or np.max(d) > 1.0:
            raise ValueError("Make sure all image data is in the range [0, 1].")
        return d


class RenderImage:
    """image/render_image.py: XYZ + power histogram of detector hits.

    `render(p, w, wl)` bins on the device (engine.render_xyzw → otb_render_xyzw); images produced by
    Raytracer.detector_image / iterative_render are created device-side and materialised lazily."""

    EPS = 1e-9
    K = 683.0   # luminous efficacy, lm/W (scipy.constants "luminous efficacy")
    SIZES = [1, 3, 5, 7, 9, 15, 21, 27, 35, 45, 63, 105, 135, 189, 315, 945]
    MAX_IMAGE_SIDE = SIZES[-1]
    MAX_IMAGE_RATIO = SIZES[2]
    image_modes = ["sRGB (Absolute RI)", "sRGB (Perceptual RI)", "Outside sRGB Gamut", "Irradiance", "Illuminance",
                   "Lightness (CIELUV)", "Hue (CIELUV)", "Chroma (CIELUV)", "Saturation (CIELUV)"]
    _MODE_IDS = {"Irradiance": 0, "Illuminance": 1, "sRGB (Absolute RI)": 2, "sRGB (Perceptual RI)": 3,
                 "Outside sRGB Gamut": 4, "Lightness (CIELUV)": 5, "Hue (CIELUV)": 6, "Chroma (CIELUV)": 7,
                 "Saturation (CIELUV)": 8}

    def __init__(self, extent, projection: str = None, desc: str = "", long_desc: str = ""):
        e = np.array(np.asarray_chkfinite(extent, dtype=np.float64))
        if e.shape[0] != 4:
            raise ValueError("Extent needs to have 4 elements.")
        self.extent = e
        self._extent0 = e.copy()
        self._data = None
        self._data_dev = None      # torch tensor (Ny, Nx, 4) float64 on the GPU
        self._counts_dev = None    # torch tensor (Ny, Nx) int32: ray counts per bin (parity checks)
        self._ready = None         # CUDA event: device image complete (side-stream all-reduce of the shards)
        self._pack = None          # engine.TilePack: occupied tiles of the image (sparse all-reduce / download)
        self._lib = None           # engine library that rendered the image (tile kernels)
        self.transferred_bytes = 0  # device -> host bytes the last materialisation moved
        self._counts_local = False  # several GPUs: the count channel still holds this rank's shard only
        self._host_buf = None      # pinned staging tensor of an asynchronous download
        self._host_ready = None    # CUDA event: download complete
        self._limit = None
        self.projection = projection
        self.desc, self.long_desc = desc, long_desc

    # -- geometry of the pixel grid (render_image.py:224-254, 383-387) ------------------------------
    @property
    def s(self):
        return [float(self.extent[1] - self.extent[0]), float(self.extent[3] - self.extent[2])]

    def _fix_extent(self) -> None:
        sx, sy = self.s
        MR = self.MAX_IMAGE_RATIO
        self.extent = self._extent0.copy()
        if sx < 2*self.EPS and sy < 2*self.EPS:
            self.extent += self.EPS*np.array([-1, 1, -1, 1])
        elif not sx or sy/sx > MR:
            xm = (self._extent0[0] + self._extent0[1])/2
            self.extent[0] = xm - sy/MR/2
            self.extent[1] = xm + sy/MR/2
        elif not sy or sx/sy > MR:
            ym = (self._extent0[2] + self._extent0[3])/2
            self.extent[2] = ym - sx/MR/2
            self.extent[3] = ym + sx/MR/2
        if self._limit is not None:
            self.extent += np.array([-1.0, 1.0, -1.0, 1.0])*2.7*self._limit/1000.0

    def _grid(self):
        Nrs = self.MAX_IMAGE_SIDE
        nf = lambda a: min(self.MAX_IMAGE_RATIO, 1 + 2*int(a/2))
        Nx = Nrs if self.s[0] <= self.s[1] else Nrs*nf(self.s[0]/self.s[1])
        Ny = Nrs if self.s[0] > self.s[1] else Nrs*nf(self.s[1]/self.s[0])
        return Nx, Ny

    # -- data access ----------------------------------------------------------------------------
    def has_image(self) -> bool:
        return self._data is not None or self._data_dev is not None

    def _wait_device(self):
        """make the current stream wait for the (side-stream) completion of the device image.  A sparse all-reduce
        (dist.allreduce_image_async) that met more occupied tiles than its learnt capacity left the image unreduced:
        detected here from the pack header and repaired with the dense all-reduce (collective: every rank sees the
        same union count and takes the same branch)."""
        if self._ready is not None:
            import torch
            torch.cuda.current_stream().wait_event(self._ready)
            self._ready = None
            if self._pack is not None and self._pack.reduced:
                from . import dist, engine
                h = self._pack.header[:2].cpu()          # synchronises: the image is about to be consumed anyway
                engine._tile_cap[(self._pack.Ny, self._pack.Nx)] = int(h[0])
                if int(h[1]):
                    dist.allreduce_sum_(self._data_dev)
                    self._pack = None

    def download_async(self) -> "RenderImage":
        """Extension of the reference API: start the device -> host copy of the image on the side stream (pinned
        staging buffer) and return at once; `data` / `_materialise()` complete it.  Lets the next trace overlap
        the transfer (a 4725 x 945 x 4 float64 image is 143 MB)."""
        if self._data is None and self._data_dev is not None and self._host_buf is None:
            import torch
            from . import engine
            side = engine.side_stream()
            if self._ready is not None:
                side.wait_event(self._ready)
            else:
                side.wait_stream(torch.cuda.current_stream())
            tp = self._pack
            cap = tp.cap if tp is not None else (engine.tile_capacity(self._data_dev.shape) if self._lib is not None else 0)
            with torch.cuda.stream(side):
                if cap:
                    # only the occupied tiles travel: [count, overflow, ids] + cap tiles instead of the whole histogram
                    if tp is None:
                        tp = engine.TilePack(self._lib, self._data_dev, cap)
                        tp.make_mask()
                        tp.pack()
                        self._pack = tp
                    self._host_hdr = engine.pinned_take(tp.header.shape, tp.header.dtype)
                    self._host_buf = engine.pinned_take(tp.packed.shape, tp.packed.dtype)
                    self._host_hdr.copy_(tp.header, non_blocking=True)
                    self._host_buf.copy_(tp.packed, non_blocking=True)
                else:
                    self._host_hdr = None
                    self._host_buf = engine.pinned_take(self._data_dev.shape, self._data_dev.dtype)
                    self._host_buf.copy_(self._data_dev, non_blocking=True)
                self._data_dev.record_stream(side)
                self._host_ready = torch.cuda.Event()
                self._host_ready.record(side)
            if cap:
                # the host-side scatter of the tiles into the dense image runs on a worker thread as soon as the copy
                # has landed (the caller is usually blocked in the next trace's status readback by then)
                self._bg = engine.assemble_submit(tuple(self._data_dev.shape), self._host_ready, self._host_hdr, self._host_buf)
        return self

    def _materialise(self):
        if self._data is None:
            if self._data_dev is None:
                raise RuntimeError("Image was not calculated/rendered yet.")
            if self._host_buf is not None:
                bg = self.__dict__.pop("_bg", None)
                done = bg.result() if bg is not None else None       # (dense view, pool token) or None (overflow)
                self._host_ready.synchronize()
                hdr = getattr(self, "_host_hdr", None)
                if hdr is None:
                    self._data = self._host_buf.numpy()       # zero-copy view of the pinned buffer
                    self.transferred_bytes = self._data.nbytes
                else:
                    from . import engine
                    h = hdr.numpy()
                    engine._tile_cap[(int(self._data_dev.shape[0]), int(self._data_dev.shape[1]))] = int(h[0])
                    if int(h[1]):                              # more tiles than the learnt capacity: dense copy
                        self._wait_device()
                        self._data = self._data_dev.cpu().numpy()
                        self.transferred_bytes = self._data.nbytes + h.nbytes + self._host_buf.numel()*8
                    else:
                        if done is None:
                            done = engine.assemble_tiles(self._data_dev.shape, h, self._host_buf.numpy())
                        self._data, self._dense_token = done
                        self.transferred_bytes = h.nbytes + self._host_buf.numel()*8
                    engine.pinned_give(self._host_buf)
                    engine.pinned_give(hdr)
                    self._host_buf = self._host_hdr = None
            else:
                self._wait_device()
                self._data = self._data_dev.cpu().numpy()
                self.transferred_bytes = self._data.nbytes
        return self._data

    def __del__(self):
        try:
            bg = self.__dict__.pop("_bg", None)
            if bg is not None:                 # download started but never consumed: take the result back
                done = bg.result()
                if done is not None:
                    self._dense_token = done[1]
            tok = self.__dict__.get("_dense_token")
            if tok is not None:
                from . import engine
                self._data = None
                self._dense_token = None
                engine.release_dense_async(tok)      # clearing the written tiles: worker thread
            if self._host_buf is not None:
                from . import engine
                if self._host_ready is not None:
                    self._host_ready.synchronize()
                self._data = None
                engine.pinned_give(self._host_buf)
                self._host_buf = None
        except Exception:
            pass

    @property
    def shape(self):
        if self._data_dev is not None:
            return tuple(self._data_dev.shape)
        return self._materialise().shape

    @property
    def data(self) -> np.ndarray:
        return self._materialise().copy()

    @property
    def counts(self) -> np.ndarray:
        """number of rays binned per pixel (int32), for count-exact parity checks.  On several GPUs the first access
        all-reduces the channel: a collective call, every rank has to make it."""
        if self._counts_dev is None:
            raise RuntimeError("No count channel rendered.")
        self._wait_device()
        if self._counts_local:
            from . import dist
            dist.allreduce_sum_(self._counts_dev)
            self._counts_local = False
        return self._counts_dev.cpu().numpy()

    @property
    def Apx(self) -> float:
        return self.s[0]*self.s[1]/(self.shape[1]*self.shape[0])

    @property
    def limit(self):
        return self._limit

    def power(self) -> float:
        if self._data is None and self._data_dev is not None:
            self._wait_device()
            return float(self._data_dev[:, :, 3].sum().item())
        return float(np.sum(self._materialise()[:, :, 3]))

    def luminous_power(self) -> float:
        if self._data is None and self._data_dev is not None:
            self._wait_device()
            return float(self.K*self._data_dev[:, :, 1].sum().item())
        return float(self.K*np.sum(self._materialise()[:, :, 1]))

    def get(self, mode: str, N: int = 315, L_th: float = 0, chroma_scale: float = None):
        """RenderImage.get (render_image.py:131-222): converted image of mode `mode`, rescaled by joining bins to
        N pixels on the smaller side (nearest of SIZES).  Rescaling and all per-pixel conversions run on the
        device (engine.image_get -> otb_image_rescale / otb_image_stats / otb_image_convert); the result is an
        RGBImage (sRGB modes) or a ScalarImage like in the reference."""
        from . import engine
        if not self.has_image():
            raise RuntimeError("Image was not calculated/rendered yet.")
        N = int(N)
        if not 1 <= N <= self.MAX_IMAGE_SIDE:
            raise ValueError(f"N needs to be between 1 and {self.MAX_IMAGE_SIDE}")
        if mode not in self._MODE_IDS:
            raise ValueError(f"Invalid display_mode {mode}, should be one of {self.image_modes}.")
        iargs = dict(extent=self.extent, projection=self.projection, desc=self.desc, long_desc=self.long_desc,
                     quantity=mode, limit=self.limit)
        Na = self.SIZES[int(np.argmin(np.abs(N - np.array(self.SIZES))))]
        fact = int(self.MAX_IMAGE_SIDE/Na)
        if self._data_dev is None:                       # image rendered elsewhere and assigned from the host
            import torch
            self._data_dev = torch.from_numpy(np.ascontiguousarray(self._data)).to(engine.device())
        self._wait_device()
        scale = {"Irradiance": 1/self.Apx, "Illuminance": self.K/self.Apx}.get(mode, 0.0)
        data = engine.image_get(self._data_dev, fact, self._MODE_IDS[mode], scale, float(L_th), chroma_scale)
        return RGBImage(data, **iargs) if data.ndim == 3 else ScalarImage(data, **iargs)

    def render(self, p: np.ndarray = None, w: np.ndarray = None, wl: np.ndarray = None,
               limit: float = None, _dont_filter: bool = False) -> None:
        """render_image.py:361-421 with the scatter-add (and the resolution filter) done by the CUDA engine."""
        from . import engine
        self._limit = limit
        self._fix_extent()
        Nx, Ny = self._grid()
        self._data = None
        self._data_dev, self._counts_dev = engine.render_xyzw_host(p, w, wl, self.extent, Nx, Ny)
        if not _dont_filter and self._limit is not None:
            self._apply_rayleigh_filter()

    def _airy_psf(self) -> np.ndarray:
        """Airy-disc kernel of the resolution filter, render_image.py:261-280 verbatim in numpy/scipy: host setup of
        a (2 ps + 1)^2 table, not per-ray work"""
        import scipy.special
        Ny, Nx = self.shape[:2]
        px = self._limit/1000.0/(self.s[0]/Nx)
        py = self._limit/1000.0/(self.s[1]/Ny)
        ps = int(np.ceil(2.7*max(px, py)))
        ps = ps + 1 if ps % 2 else ps
        Y, X = np.mgrid[-ps:ps:(2*ps + 1)*1j, -ps:ps:(2*ps + 1)*1j]
        R = np.sqrt((X/px)**2 + (Y/py)**2)*3.8317
        Rnz = R[R != 0]
        psf = np.ones((2*ps + 1, 2*ps + 1), dtype=np.float64)
        psf[R != 0] = (2*scipy.special.j1(Rnz)/Rnz)**2
        psf[R > 10.1735] = 0
        psf *= 1/psf.sum()
        return psf

    def _apply_rayleigh_filter(self) -> None:
        """RenderImage._apply_rayleigh_filter (render_image.py:255-296): convolution of the XYZW histogram with the
        Airy-disc kernel on the device (otb_image_convolve)"""
        from . import engine
        if self._limit is not None and self.projection is not None:
            raise RuntimeError("Resolution limit filter is not applicable for a projected image.")
        if not self.has_image():
            raise RuntimeError("Image was not calculated/rendered yet.")
        if self._data_dev is None:
            import torch
            self._data_dev = torch.from_numpy(np.ascontiguousarray(self._data)).to(engine.device())
        self._wait_device()
        self._data_dev = engine.image_convolve(self._data_dev, self._airy_psf())
        self._data = None
        self._pack = None          # the tile set of the unfiltered image no longer describes it
