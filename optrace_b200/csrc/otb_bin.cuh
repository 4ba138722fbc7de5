// otb_bin.cuh — XYZW histogram accumulation shared by the detector-render kernel and the fused trace+render kernel.
#pragma once
#include "otb_common.cuh"
#include "otb_media.cuh"

// ---- histogram -----------------------------------------------------------------------------------
struct BinGrid {
    double e0, e1, e2, e3;   // extent
    double fx, fy;           // Nx / sx, Ny / sy  (misc.py:82-83)
    int Nx, Ny;
};

// misc.binning_indices_2d (misc.py:59-91): false when the position falls outside the grid
__device__ __forceinline__ bool bin_index(const BinGrid& g, double x, double y, int& xi, int& yi)
{
    const double fxv = floor(g.fx*(x - g.e0)), fyv = floor(g.fy*(y - g.e2));
    if (!(fxv >= -1.0) || !(fyv >= -1.0) || !(fxv <= (double)g.Nx) || !(fyv <= (double)g.Ny)) return false;  // incl. NaN
    xi = (int)fxv;
    yi = (int)fyv;
    if (y == g.e3) yi = g.Ny - 1;
    if (x == g.e1) xi = g.Nx - 1;
    return !((xi < 0) || (yi < 0) || (yi >= g.Ny) || (xi >= g.Nx));
}

// Warp-level aggregation (must be called by all 32 lanes; `ok` marks lanes that carry a hit).
// Lanes whose hits fall into the same pixel are found with __match_any_sync, their four channel values are
// summed by a rank-ordered tree reduction inside the peer group.  Returns true on the group leaders, which then
// hold the pixel index, the channel sums and the hit count of their group.  PSF-like images (double Gauss: 7 M
// hits in 3e4 pixels, 2e5 in the hottest one) otherwise serialise on a handful of addresses; spread images skip
// the reduction after one vote.
__device__ __forceinline__ bool aggregate_xyz_warp(const BinGrid& g, bool ok, double x, double y, float w,
                                                   double ox, double oy, double oz, unsigned lane,
                                                   double& v0, double& v1, double& v2, double& v3, int& n, int& pix)
{
    int xi = 0, yi = 0;
    ok = ok && bin_index(g, x, y, xi, yi);
    v0 = v1 = v2 = v3 = 0.0;
    if (ok) {
        const double wd = (double)w;
        v0 = ox*wd;
        v1 = oy*wd;
        v2 = oz*wd;
        v3 = wd;
    }
    pix = ok ? yi*g.Nx + xi : -1 - (int)lane;          // lanes without a hit never share a key
    const unsigned peers = __match_any_sync(0xffffffffu, pix);
    const int size = __popc(peers);
    n = ok ? 1 : 0;
    if (__any_sync(0xffffffffu, size > 1)) {
        const int rank = __popc(peers & ((1u << lane) - 1));
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            // partner = lane holding rank + s of my group (if any); every lane shuffles, only even multiples add
            const int want = rank + s;
            const int src = (want < size) ? (int)__fns(peers, 0, want + 1) : (int)lane;
            const double a0 = __shfl_sync(0xffffffffu, v0, src), a1 = __shfl_sync(0xffffffffu, v1, src);
            const double a2 = __shfl_sync(0xffffffffu, v2, src), a3 = __shfl_sync(0xffffffffu, v3, src);
            const int an = __shfl_sync(0xffffffffu, n, src);
            if (want < size && (rank & (2*s - 1)) == 0) {
                v0 += a0;
                v1 += a1;
                v2 += a2;
                v3 += a3;
                n += an;
            }
        }
        ok = ok && rank == 0;
    }
    return ok;
}

// aggregation + fp64 red.global.add atomics by the group leaders
__device__ __forceinline__ void accumulate_xyz_warp(const BinGrid& g, bool ok, double x, double y, float w,
                                                    double ox, double oy, double oz, double* __restrict__ img, int* __restrict__ cnt)
{
    double v0, v1, v2, v3;
    int n, pix;
    ok = aggregate_xyz_warp(g, ok, x, y, w, ox, oy, oz, threadIdx.x & 31, v0, v1, v2, v3, n, pix);
    if (ok) {
        double* q = img + 4*(int64_t)pix;
        atomicAdd(q + 0, v0);
        atomicAdd(q + 1, v1);
        atomicAdd(q + 2, v2);
        atomicAdd(q + 3, v3);
        if (cnt) atomicAdd(cnt + pix, n);
    }
}

// same with the CIE observer lookup for the hit's wavelength
__device__ __forceinline__ void accumulate_hit_warp(const BinGrid& g, const double* __restrict__ obs, bool ok, double x, double y,
                                                    float w, float wl, double* __restrict__ img, int* __restrict__ cnt)
{
    double ox = 0.0, oy = 0.0, oz = 0.0;
    if (ok) observer_xyz(obs, (double)wl, ox, oy, oz);
    accumulate_xyz_warp(g, ok, x, y, w, ox, oy, oz, img, cnt);
}
