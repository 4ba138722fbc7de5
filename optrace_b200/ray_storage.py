"""Device-resident ray storage with lazy host materialisation.

Contract of optrace/tracer/ray_storage.py (array names, dtypes, shapes, Fortran order, section semantics):
    p_list  float64 (N, nt, 3)    s0_list float64 (N, 3)   pol_list float32 (N, nt, 3) (NaN broadcast if no_pol)
    w_list  float32 (N, nt)       n_list  float64 (N, nt)   wl_list  float32 (N)
The arrays live on the GPU as SoA planes written by the trace kernel in exactly this byte layout, so a host
view is a plain device->host copy plus a zero-copy Fortran reshape.  Copies happen per attribute on first
access (SURVEY.md hard part 5: 48 B/ray/section over PCIe is the bottleneck, so nothing is copied eagerly).
In multi-GPU runs every rank holds a share of every source's block (dist.shard_sources), stored in source order.
"""
from __future__ import annotations

import numpy as np


def split_rays(N: int, powers) -> np.ndarray:
    """rays per source (ray_storage.py:56-68): floor share by power, remainder drawn with np.random.choice"""
    P = np.asarray(powers, dtype=np.float64)
    P_all = np.sum(P)
    N_list = (N*P/P_all).astype(int)
    dN = N - np.sum(N_list)
    idx = np.random.choice(N_list.shape[0], size=dN, p=P/P_all)
    np.add.at(N_list, idx, np.ones(idx.shape, dtype=int))
    return N_list


class RayStorage:

    def __init__(self):
        self.N_list = np.array([], dtype=int)
        self.B_list = np.array([], dtype=int)
        self.no_pol = False
        self.ray_source_list = []
        self._dev = None          # engine.DeviceStore of the local shard
        self._host = {}
        self._nt = 0
        self._N_global = 0
        self.ray_begin, self.ray_end = 0, 0
        self._blocks = []         # local blocks: (source index, global id of the first ray, count), in source order

    # -- bookkeeping ----------------------------------------------------------------------------
    def _attach(self, dev_store, sources, N_list, no_pol, N_global, blocks):
        """blocks: the local rays as (source index, global first ray id, count) in storage order — one block per
        source; an int is accepted for a contiguous global range starting there"""
        self._dev = dev_store
        self._host = {}
        self.ray_source_list = list(sources)
        self.N_list = np.asarray(N_list, dtype=int)
        self.B_list = np.concatenate(([0], np.cumsum(self.N_list))).astype(int)
        self.no_pol = no_pol
        self._nt = dev_store.nt
        self._N_global = int(N_global)
        if isinstance(blocks, (int, np.integer)):
            from . import dist
            blocks = dist.contiguous_blocks(self.N_list, int(blocks), int(blocks) + dev_store.N)
        self._blocks = [(int(i), int(g), int(c)) for i, g, c in blocks]
        assert sum(c for _, _, c in self._blocks) == dev_store.N
        # first global ray id of the shard (meaningful for contiguous shards: injected bundles, single GPU)
        self.ray_begin = self._blocks[0][1] if self._blocks else 0
        self.ray_end = self.ray_begin + dev_store.N

    @staticmethod
    def storage_size(N: int, nt: int, no_pol: bool) -> int:
        """ray_storage.py:92-104"""
        fpol = 4*N*nt*3 if not no_pol else 8
        return N*nt*3*8 + N*3*8 + fpol + N*nt*4 + N*nt*8 + N*4

    @staticmethod
    def max_rays_for_size(size: int, nt: int, no_pol: bool) -> int:
        """ray_storage.py:106-122"""
        if no_pol:
            return (size - 8)//(nt*3*8 + 3*8 + nt*4 + nt*8 + 4)
        return size//(nt*3*8 + 3*8 + 4*nt*3 + nt*4 + nt*8 + 4)

    @property
    def N(self) -> int:
        """number of rays held by this process (the local shard in multi-GPU runs)"""
        return self._dev.N if self._dev is not None and self.N_list.shape[0] else 0

    @property
    def N_global(self) -> int:
        return self._N_global

    @property
    def Nt(self) -> int:
        return self._nt if self.N_list.shape[0] else 0

    def crepr(self):
        return [tuple(self.N_list), tuple(self.B_list), self.no_pol, id(self._dev)]

    # -- lazy host views -------------------------------------------------------------------------
    def _fetch(self, key, tensor, shape):
        if key not in self._host:
            if self._dev is None:
                return np.array([])
            a = tensor.cpu().numpy().reshape(shape, order="F")
            a.flags.writeable = False
            self._host[key] = a
        return self._host[key]

    @property
    def p_list(self) -> np.ndarray:
        return self._fetch("p", self._dev.p, (self._dev.N, self._nt, 3)) if self._dev else np.array([])

    @property
    def s0_list(self) -> np.ndarray:
        """final directions after the trace — the reference mutates its s0_list view in place
        (ray_storage.py:170, raytracer.py:829; SURVEY.md hard part 10)"""
        return self._fetch("s", self._dev.s, (self._dev.N, 3)) if self._dev else np.array([])

    @property
    def w_list(self) -> np.ndarray:
        return self._fetch("w", self._dev.w, (self._dev.N, self._nt)) if self._dev else np.array([])

    @property
    def n_list(self) -> np.ndarray:
        return self._fetch("n", self._dev.n, (self._dev.N, self._nt)) if self._dev else np.array([])

    @property
    def wl_list(self) -> np.ndarray:
        return self._fetch("wl", self._dev.wl, (self._dev.N,)) if self._dev else np.array([])

    @property
    def pol_list(self) -> np.ndarray:
        if self._dev is None:
            return np.array([])
        if self.no_pol:
            return np.broadcast_to(np.nan, (self._dev.N, self._nt, 3))
        return self._fetch("pol", self._dev.pol, (self._dev.N, self._nt, 3))

    # -- accessors of the reference (host post-processing on the materialised arrays) ---------------
    def source_sections(self, index: int = None):
        """ray_storage.py:173-187"""
        assert self.N, "ray_source_list has no rays stored."
        Ns, Ne = self._local_range(index)
        return self.p_list[Ns:Ne, 0], self.s0_list[Ns:Ne], self.pol_list[Ns:Ne, 0], \
            self.w_list[Ns:Ne, 0], self.wl_list[Ns:Ne]

    def _local_range(self, index):
        """local storage rows [b, e) of source `index` (all local rows for None); empty when this rank holds none"""
        if index is None:
            return 0, self.N
        off = 0
        for i, _, c in self._blocks:
            if i == index:
                return off, off + c
            if i > index:
                break
            off += c
        return off, off

    def rays_by_mask(self, ch=None, ch2=None, ret=None, normalize: bool = True):
        """ray_storage.py:235-293"""
        assert self.N, "ray_source_list has no rays stored."
        ret = [1, 1, 1, 1, 1, 1, 1] if ret is None else ret
        ch = np.ones(self.N, dtype=bool) if ch is None else ch
        ch2 = slice(None) if ch2 is None else ch2
        assert ch.shape[0] == self.N
        snums = s = None
        if ret[5]:
            ind = np.nonzero(ch)[0]
            lb = np.cumsum([0] + [c for _, _, c in self._blocks])
            src = np.array([i for i, _, _ in self._blocks], dtype=int)
            snums = src[np.clip(np.searchsorted(lb, ind, side="right") - 1, 0, len(src) - 1)] if len(src) else ind*0
        if ret[1]:
            P = self.p_list
            if not isinstance(ch2, slice):
                ch21 = np.where(ch2 < self.Nt - 1, ch2 + 1, ch2)
                s = P[ch, ch21] - P[ch, ch2]
                if normalize:
                    with np.errstate(invalid="ignore"):
                        s = s/np.sqrt(s[:, 0]**2 + s[:, 1]**2 + s[:, 2]**2)[:, None]
            else:
                s = P[ch, 1:] - P[ch, :-1]
                s = np.hstack((s, np.zeros((s.shape[0], 1, 3), dtype=np.float64)))
                if normalize:
                    with np.errstate(invalid="ignore"):
                        s = s/np.sqrt(s[..., 0]**2 + s[..., 1]**2 + s[..., 2]**2)[..., None]
        return (self.p_list[ch, ch2] if ret[0] else None, s if ret[1] else None,
                self.pol_list[ch, ch2] if ret[2] else None, self.w_list[ch, ch2] if ret[3] else None,
                self.wl_list[ch] if ret[4] else None, snums if ret[5] else None,
                self.n_list[ch, ch2] if ret[6] else None)

    def ray_lengths(self, ch=None, ch2=None) -> np.ndarray:
        _, s, *_ = self.rays_by_mask(ch, ch2, ret=[0, 1, 0, 0, 0, 0, 0], normalize=False)
        return np.linalg.norm(s, axis=s.ndim - 1)

    def optical_lengths(self, ch=None, ch2=None) -> np.ndarray:
        _, s, _, _, _, _, n = self.rays_by_mask(ch, ch2, ret=[0, 1, 0, 0, 0, 0, 1], normalize=False)
        return np.linalg.norm(s, axis=s.ndim - 1)*n

    def source_numbers(self) -> np.ndarray:
        return self.rays_by_mask(ret=[0, 0, 0, 0, 0, 1, 0])[5]

    def direction_vectors(self, normalize: bool = True) -> np.ndarray:
        return self.rays_by_mask(ret=[0, 1, 0, 0, 0, 0, 0], normalize=normalize)[1]
