// otb_focus.cu — device side of Raytracer.focus_search (raytracer.py:1354-1640): the per-ray work of the focus
// cost functions on device-resident ray storage.  The optimiser (scipy) and the O(N_px^2) arithmetic on the small
// cost images stay on the host, exactly as in the reference; what runs here is everything that touches N rays:
//   * otb_focus_prepare: section selection and the auxiliary line hit(z) = pa + sb z of every ray
//     (raytracer.py:1553-1578, RayStorage.rays_by_mask ray_storage.py:235-293),
//   * otb_focus_moments: weighted sums for np.cov / np.average and the direct RMS solution
//     (raytracer.py:1376-1379, 1424-1447, 1620),
//   * otb_focus_image: hit range + weighted N_px x N_px histogram (raytracer.py:1387-1392,
//     misc.binning_indices_2d misc.py:59-91).
// All kernels are streaming passes over SoA arrays (HBM-bound); reductions are per-block in shared memory with one
// fp64 atomic per block and quantity.
#include "otb_common.cuh"
#include "otb_bin.cuh"

__device__ __forceinline__ void fatomic_min(double* addr, double v)
{
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (!(v < __longlong_as_double((long long)assumed))) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void fatomic_max(double* addr, double v)
{
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (!(v > __longlong_as_double((long long)assumed))) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

// block sum of NS doubles -> one atomicAdd per quantity
template <int NS>
__device__ __forceinline__ void block_sum_store(double (&v)[NS], double* out)
{
    __shared__ double sm[NS][8];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < NS; ++k) sm[k][warp] = v[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < NS; ++k) {
            double r = sm[k][0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r += sm[k][w];
            atomicAdd(&out[k], r);
        }
}

struct FocusLines {      // per selected ray: hit(z) = (pax + sbx z, pay + sby z), weight, used flag
    double *pax, *pay, *sbx, *sby;
    float* w;
    unsigned char* use;
};

__global__ void __launch_bounds__(256) focus_prepare_kernel(const OtbRayStore st, int64_t begin, int64_t end, double z,
                                                            FocusLines L, unsigned long long* n_use)
{
    const int64_t N = st.N;
    const int nt = st.nt;
    const int64_t Nnt = N*(int64_t)nt;
    const int64_t ray = begin + (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    bool used = false;
    if (ray < end) {
        const double* __restrict__ P = st.p_d;
        // pos = np.argmax(z < p_list[:, :, 2], axis=1) - 1: section before the first point behind z
        int j = -1;
        for (int k = 0; k < nt; ++k)
            if (z < P[ray + N*(int64_t)k + 2*Nnt]) { j = k; break; }
        const int pos = (j < 0 ? 0 : j) - 1;            // no point behind z: argmax = 0 -> pos = -1 (excluded)
        const int64_t i = ray - begin;
        if (pos >= 0) {
            const int p1 = (pos < nt - 1) ? pos + 1 : pos;
            const int64_t o0 = ray + N*(int64_t)pos, o1 = ray + N*(int64_t)p1;
            const V3 p = v3(P[o0], P[o0 + Nnt], P[o0 + 2*Nnt]);
            const V3 s = unit3(v3(P[o1] - p.x, P[o1 + Nnt] - p.y, P[o1 + 2*Nnt] - p.z));     // misc.normalize
            // pa = p - s/s_z*p_z ; sb = s/s_z   (raytracer.py:1575-1576)
            const double sx = s.x/s.z, sy = s.y/s.z;
            L.pax[i] = p.x - sx*p.z;
            L.pay[i] = p.y - sy*p.z;
            L.sbx[i] = sx;
            L.sby[i] = sy;
            L.w[i] = st.w_d[o0];
            used = true;
        } else {
            L.pax[i] = L.pay[i] = L.sbx[i] = L.sby[i] = 0.0;
            L.w[i] = 0.0f;
        }
        L.use[i] = used ? 1 : 0;
    }
    const unsigned b = __ballot_sync(0xffffffffu, used);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_use, (unsigned long long)__popc(b));
}

// mode 0: out = [sum w, sum w^2, sum w x, sum w y]                                  at z = par[0]
// mode 1: out = [sum w (x - par[1])^2, sum w (y - par[2])^2]                        at z = par[0]
// mode 2: out = [sum w^2 dtx^2 + w^2 dty^2, sum dtx dx w^2 + dty dy w^2] with dx = pax - par[1], dy = pay - par[2],
//         dtx = sbx - par[3], dty = sby - par[4]                                   (raytracer.py:1430-1442)
__global__ void __launch_bounds__(256) focus_moments_kernel(FocusLines L, int64_t n, int mode, const double* __restrict__ par,
                                                            double* out)
{
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const double z = par[0];
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x) {
        if (!L.use[i]) continue;
        const double w = (double)L.w[i];
        if (mode == 0) {
            const double x = L.pax[i] + L.sbx[i]*z, y = L.pay[i] + L.sby[i]*z;
            v[0] += w;
            v[1] += w*w;
            v[2] += w*x;
            v[3] += w*y;
        } else if (mode == 1) {
            const double x = L.pax[i] + L.sbx[i]*z, y = L.pay[i] + L.sby[i]*z;
            const double dx = x - par[1], dy = y - par[2];
            v[0] += w*(dx*dx);
            v[1] += w*(dy*dy);
        } else {
            const double w2 = w*w;
            const double dx = L.pax[i] - par[1], dy = L.pay[i] - par[2];
            const double dtx = L.sbx[i] - par[3], dty = L.sby[i] - par[4];
            v[0] += w2*(dtx*dtx) + w2*(dty*dty);
            v[1] += dtx*dx*w2 + dty*dy*w2;
        }
    }
    block_sum_store<4>(v, out);
}

// range of the hit positions at z over the used rays: rng = [min x, max x, min y, max y]
__global__ void __launch_bounds__(256) focus_range_kernel(FocusLines L, int64_t n, double z, double* rng)
{
    double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x) {
        if (!L.use[i]) continue;
        const double x = L.pax[i] + L.sbx[i]*z, y = L.pay[i] + L.sby[i]*z;
        mnx = fmin(mnx, x); mxx = fmax(mxx, x);
        mny = fmin(mny, y); mxy = fmax(mxy, y);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
        mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, d));
        mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
    }
    if ((threadIdx.x & 31) == 0 && mnx <= mxx) {
        fatomic_min(&rng[0], mnx);
        fatomic_max(&rng[1], mxx);
        fatomic_min(&rng[2], mny);
        fatomic_max(&rng[3], mxy);
    }
}

// weighted N_px x N_px histogram of the hit positions at z; the grid extent is read from device memory (the range
// pass above), so no host round trip is needed between the two passes
__global__ void __launch_bounds__(256) focus_image_kernel(FocusLines L, int64_t n, double z, const double* __restrict__ rng,
                                                          int npx, double* __restrict__ img)
{
    BinGrid g;
    g.e0 = rng[0]; g.e1 = rng[1]; g.e2 = rng[2]; g.e3 = rng[3];
    g.Nx = g.Ny = npx;
    g.fx = (double)npx/(g.e1 - g.e0);
    g.fy = (double)npx/(g.e3 - g.e2);
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x) {
        if (!L.use[i]) continue;
        const double x = L.pax[i] + L.sbx[i]*z, y = L.pay[i] + L.sby[i]*z;
        int xi, yi;
        if (bin_index(g, x, y, xi, yi)) atomicAdd(&img[(int64_t)yi*npx + xi], (double)L.w[i]);
    }
}

int otb_sm_count();

extern "C" {

int otb_focus_prepare(const OtbRayStore* store, int64_t ray_begin, int64_t ray_end, double z,
                      double* pax_d, double* pay_d, double* sbx_d, double* sby_d, float* w_d, uint8_t* use_d,
                      int64_t* n_use_d, void* stream)
{
    if (!store || !pax_d || !pay_d || !sbx_d || !sby_d || !w_d || !use_d || !n_use_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (ray_begin < 0 || ray_end > store->N || ray_begin > ray_end) { otb_set_error("invalid ray range"); return OTB_ERR_INVALID_ARG; }
    const int64_t n = ray_end - ray_begin;
    if (n == 0) return OTB_OK;
    FocusLines L = {pax_d, pay_d, sbx_d, sby_d, w_d, use_d};
    focus_prepare_kernel<<<(unsigned)((n + 255)/256), 256, 0, (cudaStream_t)stream>>>(*store, ray_begin, ray_end, z, L,
                                                                                      (unsigned long long*)n_use_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

static int focus_grid(int64_t n) { const int64_t b = (n + 255)/256, cap = 8LL*otb_sm_count(); return (int)(b < cap ? b : cap); }

int otb_focus_moments(const double* pax_d, const double* pay_d, const double* sbx_d, const double* sby_d, const float* w_d,
                      const uint8_t* use_d, int64_t n, int32_t mode, const double* par_d, double* out_d, void* stream)
{
    if (!pax_d || !pay_d || !sbx_d || !sby_d || !w_d || !use_d || !par_d || !out_d || n < 0 || mode < 0 || mode > 2) {
        otb_set_error("invalid argument");
        return OTB_ERR_INVALID_ARG;
    }
    FocusLines L = {(double*)pax_d, (double*)pay_d, (double*)sbx_d, (double*)sby_d, (float*)w_d, (unsigned char*)use_d};
    OTB_CUDA(cudaMemsetAsync(out_d, 0, 4*sizeof(double), (cudaStream_t)stream));
    if (n == 0) return OTB_OK;      // empty shard (multi-GPU: this rank holds no ray of the selected source): zero sums
    focus_moments_kernel<<<focus_grid(n), 256, 0, (cudaStream_t)stream>>>(L, n, mode, par_d, out_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

// phase 0: both passes; phase 1: only the hit range into rng_d[4]; phase 2: only the weighted histogram img_d
// (npx*npx doubles) over the range found in rng_d (several GPUs: the caller all-reduces the range in between)
int otb_focus_image(const double* pax_d, const double* pay_d, const double* sbx_d, const double* sby_d, const float* w_d,
                    const uint8_t* use_d, int64_t n, double z, int32_t npx, int32_t phase, double* rng_d, double* img_d,
                    void* stream)
{
    if (!pax_d || !pay_d || !sbx_d || !sby_d || !w_d || !use_d || !rng_d || !img_d || n < 0 || npx < 1 || phase < 0 || phase > 2) {
        otb_set_error("invalid argument");
        return OTB_ERR_INVALID_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    FocusLines L = {(double*)pax_d, (double*)pay_d, (double*)sbx_d, (double*)sby_d, (float*)w_d, (unsigned char*)use_d};
    // n == 0 is a valid empty shard (multi-GPU: no ray of the selected source on this rank): the range stays at its
    // neutral element, the histogram at zero, and the caller's collectives still line up across ranks
    if (phase != 2) {
        static const double init[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
        OTB_CUDA(cudaMemcpyAsync(rng_d, init, sizeof(init), cudaMemcpyHostToDevice, st));
        if (n > 0) focus_range_kernel<<<focus_grid(n), 256, 0, st>>>(L, n, z, rng_d);
    }
    if (phase != 1) {
        OTB_CUDA(cudaMemsetAsync(img_d, 0, sizeof(double)*(size_t)npx*npx, st));
        if (n > 0) focus_image_kernel<<<focus_grid(n), 256, 0, st>>>(L, n, z, rng_d, npx, img_d);
    }
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

}  // extern "C"
