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

__device__ __forceinline__ void accumulate_hit(const BinGrid& g, const double* __restrict__ obs, double x, double y,
                                               float w, float wl, double* __restrict__ img, int* __restrict__ cnt)
{
    int xi, yi;
    if (!bin_index(g, x, y, xi, yi)) return;
    double ox, oy, oz;
    observer_xyz(obs, (double)wl, ox, oy, oz);
    const double wd = (double)w;
    const int64_t pix = (int64_t)yi*g.Nx + xi;
    double* q = img + 4*pix;
    atomicAdd(q + 0, ox*wd);
    atomicAdd(q + 1, oy*wd);
    atomicAdd(q + 2, oz*wd);
    atomicAdd(q + 3, wd);
    if (cnt) atomicAdd(cnt + pix, 1);
}

