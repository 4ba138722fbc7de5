// otb_spectrum.cu — weighted wavelength histograms of device-resident rays: LightSpectrum.render
// (light_spectrum.py:40-79) behind Raytracer.detector_spectrum / source_spectrum (raytracer.py:1100-1132,
// 1311-1329).  Two passes like the reference's host code: statistics (count of non-zero weights for the bin
// number, wavelength range), then np.histogram's uniform-bin index rule on the float32 wavelengths.
#include "otb_common.cuh"

// atomic min / max on floats through their ordered integer image (values are finite wavelengths)
__device__ __forceinline__ void atomic_min_f(float* addr, float v)
{
    int* a = (int*)addr;
    int old = *a, assumed;
    do {
        assumed = old;
        if (!(v < __int_as_float(assumed))) break;
        old = atomicCAS(a, assumed, __float_as_int(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v)
{
    int* a = (int*)addr;
    int old = *a, assumed;
    do {
        assumed = old;
        if (!(v > __int_as_float(assumed))) break;
        old = atomicCAS(a, assumed, __float_as_int(v));
    } while (assumed != old);
}

// stats_d: [0] rays used (int64), [1] rays with non-zero weight among them (int64), [2] = two floats: min, max wl
__global__ void __launch_bounds__(256) spectrum_stats_kernel(const float* __restrict__ wl, const float* __restrict__ w, int64_t M,
                                                             int positive_only, unsigned long long* cnt, float* rng)
{
    float mn = INFINITY, mx = -INFINITY;
    unsigned long long used = 0, nz = 0;
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x*blockDim.x) {
        const float wi = w[i];
        if (positive_only && !(wi > 0.0f)) continue;       // detector path: only valid hits reach the spectrum
        const float l = wl[i];
        mn = fminf(mn, l);
        mx = fmaxf(mx, l);
        ++used;
        nz += (wi != 0.0f);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        used += __shfl_xor_sync(0xffffffffu, used, d);
        nz += __shfl_xor_sync(0xffffffffu, nz, d);
    }
    if ((threadIdx.x & 31) == 0 && used) {
        atomicAdd(&cnt[0], used);
        atomicAdd(&cnt[1], nz);
        atomic_min_f(&rng[0], mn);
        atomic_max_f(&rng[1], mx);
    }
}

// np.histogram with `range` and an integer bin count (numpy/lib/_histograms_impl.py, "fast algorithm for equal
// bins"): float32 data and float32 range give float32 bin edges; index = int((a - first) / (last - first) * n) in
// float32, index n folded into the last bin, then corrected against the edges on both sides.
__global__ void __launch_bounds__(256) spectrum_hist_kernel(const float* __restrict__ wl, const float* __restrict__ w, int64_t M,
                                                            int positive_only, const float* __restrict__ edges, int nbins,
                                                            int use_smem, double* __restrict__ hist)
{
    extern __shared__ double sh[];
    if (use_smem) {
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0.0;
        __syncthreads();
    }
    const float first = edges[0], last = edges[nbins];
    const float denom = __fsub_rn(last, first);
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x*blockDim.x) {
        const float wi = w[i];
        if (positive_only && !(wi > 0.0f)) continue;
        const float a = wl[i];
        if (!(a >= first) || !(a <= last)) continue;
        int k = (int)__fmul_rn(__fdiv_rn(__fsub_rn(a, first), denom), (float)nbins);
        if (k == nbins) --k;
        if (a < edges[k]) --k;
        if (a >= edges[k + 1] && k != nbins - 1) ++k;
        if (use_smem) atomicAdd(&sh[k], (double)wi);
        else atomicAdd(&hist[k], (double)wi);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += blockDim.x)
            if (sh[i] != 0.0) atomicAdd(&hist[i], sh[i]);
    }
}

int otb_sm_count();

extern "C" {

int otb_spectrum_stats(const float* wl_d, const float* w_d, int64_t M, int32_t positive_only, int64_t* count_d,
                       float* range_d, void* stream)
{
    if (!count_d || !range_d || M < 0 || (M > 0 && (!wl_d || !w_d))) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    if (M == 0) return OTB_OK;
    const int64_t b = (M + 255)/256, cap = 8LL*otb_sm_count();
    spectrum_stats_kernel<<<(int)(b < cap ? b : cap), 256, 0, (cudaStream_t)stream>>>(wl_d, w_d, M, positive_only,
                                                                                   (unsigned long long*)count_d, range_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

int otb_spectrum_hist(const float* wl_d, const float* w_d, int64_t M, int32_t positive_only, const float* edges_d,
                      int32_t nbins, double* hist_d, void* stream)
{
    if (!edges_d || !hist_d || nbins < 1 || M < 0 || (M > 0 && (!wl_d || !w_d))) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    if (M == 0) return OTB_OK;
    const size_t smem = sizeof(double)*(size_t)nbins;
    const int use_smem = smem <= 40*1024;
    const int64_t b = (M + 255)/256, cap = 4LL*otb_sm_count();
    spectrum_hist_kernel<<<(int)(b < cap ? b : cap), 256, use_smem ? smem : 0, (cudaStream_t)stream>>>(
        wl_d, w_d, M, positive_only, edges_d, nbins, use_smem, hist_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

}  // extern "C"
