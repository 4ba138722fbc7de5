// otb_detect.cu — detector path: hit finding on stored ray sections (Raytracer._hit_detector,
// raytracer.py:881-1051) and XYZW histogram binning (RenderImage.render, render_image.py:390-417,
// misc.binning_indices_2d, misc.py:59-91, CIE observers, observers.py:14-41).
#include <mutex>
#include "otb_common.cuh"
#include "otb_surfaces.cuh"
#include "otb_media.cuh"
#include "otb_observers.cuh"
#include "otb_bin.cuh"


struct DetArgs {
    OtbRayStore st;
    KSurface surf;
    int projection, has_extent;
    double extent[4];
    int64_t begin, end;
    double* hx;
    double* hy;
    float* hw;
    double* range;           // [4] min x, max x, min y, max y
    unsigned long long* ill;
    int* status;
};

#define OTB_DET_COARSE 8      // coarse samples of the two-level section search: covers nt <= 32

#ifndef OTB_DET_THREADS
#define OTB_DET_THREADS 128
#endif
#ifndef OTB_DET_RESIDENT
#define OTB_DET_RESIDENT 768      // resident threads per SM the register allocation is made for
#endif
#ifdef OTB_DET_PLAINLD
#define OTB_DET_LD(p) (*(p))
#else
#define OTB_DET_LD(p) __ldcs(p)
#endif
__global__ void __launch_bounds__(OTB_DET_THREADS, OTB_DET_RESIDENT/OTB_DET_THREADS) detector_hits_kernel(const DetArgs a)
{
    const int64_t N = a.st.N;
    const int nt = a.st.nt;
    const int64_t Nnt = N*(int64_t)nt;
    const double* __restrict__ P = a.st.p_d;
    const float* __restrict__ Wt = a.st.w_d;
    const KSurface& S = a.surf;
    const int64_t ray = a.begin + (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    const bool valid = ray < a.end;
    const bool monotone = a.st.trace_status_d && !(__ldg(a.st.trace_status_d) & OTB_STATUS_Z_DECREASE);

    float w = 0.0f;
    double X = 0.0, Y = 0.0;
    bool ill = false, ok = false;
    if (valid) {
        // section straddling the detector z-extent (raytracer.py:929-938)
        bool no_start = true, no_reach = true;
        int first_ge = -1;
        const double* __restrict__ Pz = P + ray + 2*Nnt;
        // Image planes: a detector behind every tracing surface is reached in the LAST section by every ray.  With z
        // monotone along the rays (reported by the trace) a second last point in front of z_min says so for this ray
        // with two loads instead of the search over the sections below.
        double z_prev = 0.0, z_last = 0.0;
        bool last_section = false;
        // ... and everything the hit of that section needs (both end points and the weight: 7 independent loads, the
        // same bytes the walk below would read in three DEPENDENT round trips — the kernel is bound by DRAM latency)
        // is requested in the same batch.
        V3 pa = v3(0, 0, 0), pb = v3(0, 0, 0);
        float wa = 0.0f;
        if (monotone && nt >= 2) {
            const int64_t oa = ray + N*(int64_t)(nt - 2), ob = ray + N*(int64_t)(nt - 1);
            pa = v3(OTB_DET_LD(P + oa), OTB_DET_LD(P + oa + Nnt), OTB_DET_LD(P + oa + 2*Nnt));
            pb = v3(OTB_DET_LD(P + ob), OTB_DET_LD(P + ob + Nnt), OTB_DET_LD(P + ob + 2*Nnt));
            wa = OTB_DET_LD(Wt + oa);
            z_prev = pa.z;
            z_last = pb.z;
            last_section = z_prev < S.z_min;
        }
        if (last_section) {
            no_start = false;
            no_reach = !(z_last >= S.z_min);           // z_max >= z_min: both comparisons of the scan fail together
            first_ge = (z_last >= S.z_min) ? nt - 1 : -1;
        } else if (monotone && nt <= 4*OTB_DET_COARSE) {
            // z never decreases along a ray (reported by the trace): two batches of independent loads instead of
            // a scan over all nt sections or a bisection of dependent loads (the kernel is bound by the number
            // of DRAM round trips per ray, not by bytes): every 4th section plus the last one, then the three
            // sections inside the bracket that contains the first z >= z_min
            double zc[OTB_DET_COARSE + 1];
#pragma unroll
            for (int u = 0; u < OTB_DET_COARSE; ++u) zc[u] = (4*u < nt) ? OTB_DET_LD(Pz + N*(int64_t)(4*u)) : INFINITY;
            const double zl = OTB_DET_LD(Pz + N*(int64_t)(nt - 1));
            const double z0 = zc[0];
            no_start = (z0 >= S.z_min) && (z0 >= S.z_max);
            no_reach = !(zl >= S.z_min) && !(zl >= S.z_max);
            if (z0 >= S.z_min) first_ge = 0;
            else if (zl >= S.z_min) {
                int kb = 0;                            // last coarse sample below z_min
#pragma unroll
                for (int u = 1; u < OTB_DET_COARSE; ++u) if (4*u < nt && !(zc[u] >= S.z_min)) kb = u;
                const int j0 = 4*kb;                   // z[j0] < z_min, first_ge in (j0, min(j0 + 4, nt - 1)]
                double zf[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) zf[u] = (j0 + 1 + u < nt) ? OTB_DET_LD(Pz + N*(int64_t)(j0 + 1 + u)) : INFINITY;
                first_ge = (j0 + 4 < nt - 1) ? j0 + 4 : nt - 1;
#pragma unroll
                for (int u = 2; u >= 0; --u) if (j0 + 1 + u < nt && zf[u] >= S.z_min) first_ge = j0 + 1 + u;
            }
        } else {
            // the z plane of every section is read once; four independent loads in flight per thread
            int j = 0;
            for (; j + 4 <= nt; j += 4) {
                double z[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) z[u] = OTB_DET_LD(Pz + N*(int64_t)(j + u));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool bmin = z[u] >= S.z_min, bmax = z[u] >= S.z_max;
                    no_start = no_start && (bmin && bmax);
                    no_reach = no_reach && (!bmin && !bmax);
                    if (first_ge < 0 && bmin) first_ge = j + u;
                }
            }
            for (; j < nt; ++j) {
                const double z = OTB_DET_LD(Pz + N*(int64_t)j);
                const bool bmin = z >= S.z_min, bmax = z >= S.z_max;
                no_start = no_start && (bmin && bmax);
                no_reach = no_reach && (!bmin && !bmax);
                if (first_ge < 0 && bmin) first_ge = j;
            }
        }
        if (!(no_start || no_reach)) {
            int sec = (first_ge < 0 ? 0 : first_ge) - 1;
            if (sec < 0) sec = 0;
            HitResult h;
            h.hit = false;
            h.p = v3(0, 0, 0);
            bool more = true;
            while (more) {
                // rays_by_mask (ray_storage.py:235-293): direction from position differences, normalised
                const int s1 = (sec < nt - 1) ? sec + 1 : sec;
                const int64_t o0 = ray + N*(int64_t)sec, o1 = ray + N*(int64_t)s1;
                const bool pre = last_section && sec == nt - 2;        // the preloaded section
                const V3 p = pre ? pa : v3(P[o0], P[o0 + Nnt], P[o0 + 2*Nnt]);
                const V3 q = pre ? pb : v3(P[o1], P[o1 + Nnt], P[o1 + 2*Nnt]);
                const V3 s = unit3(v3(q.x - p.x, q.y - p.y, q.z - p.z));
                w = pre ? wa : Wt[o0];
                ++sec;
                if (sec >= nt) {          // ray ends at the outline: no intersection (raytracer.py:970-978)
                    w = 0.0f;
                    break;
                }
                h = surf_find_hit<OTB_CAPS_DET>(S, nullptr, p, s, a.status);
                ill = ill || h.ill;
                const double p2z = pre ? pb.z : P[ray + N*(int64_t)sec + 2*Nnt];
                more = h.p.z > p2z + OTB_C_EPS;      // hit behind the next stored point -> try next section
            }
            if (h.hit && w > 0.0f) {
                X = h.p.x;
                Y = h.p.y;
                sphere_project(S, a.projection, X, Y, h.p.z);
                ok = true;
                if (a.has_extent) {
                    const double* e = a.extent;
                    ok = (e[0] <= X) && (X <= e[1]) && (e[2] <= Y) && (Y <= e[3]);
                }
            }
        }
        const int64_t k = ray - a.begin;
        a.hx[k] = X;
        a.hy[k] = Y;
        a.hw[k] = ok ? w : 0.0f;
    }

    // block reduction of the hit range (auto extent, raytracer.py:1042-1046) and the ill counter
    __shared__ double smin_x[OTB_DET_THREADS/32], smax_x[OTB_DET_THREADS/32], smin_y[OTB_DET_THREADS/32], smax_y[OTB_DET_THREADS/32];
    __shared__ int sill;
    if (threadIdx.x == 0) sill = 0;
    double mnx = ok ? X : INFINITY, mxx = ok ? X : -INFINITY, mny = ok ? Y : INFINITY, mxy = ok ? Y : -INFINITY;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
        mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, d));
        mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { smin_x[warp] = mnx; smax_x[warp] = mxx; smin_y[warp] = mny; smax_y[warp] = mxy; }
    __syncthreads();
    unsigned bi = __ballot_sync(0xffffffffu, ill);
    if (bi && (threadIdx.x & 31) == 0) atomicAdd(&sill, __popc(bi));
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            mnx = fmin(smin_x[0], smin_x[k]); smin_x[0] = mnx;
            mxx = fmax(smax_x[0], smax_x[k]); smax_x[0] = mxx;
            mny = fmin(smin_y[0], smin_y[k]); smin_y[0] = mny;
            mxy = fmax(smax_y[0], smax_y[k]); smax_y[0] = mxy;
        }
#ifndef OTB_DET_NORANGE
        if (smin_x[0] <= smax_x[0]) {
            atomic_min_double(&a.range[0], smin_x[0]);
            atomic_max_double(&a.range[1], smax_x[0]);
            atomic_min_double(&a.range[2], smin_y[0]);
            atomic_max_double(&a.range[3], smax_y[0]);
        }
#endif
        if (sill) atomicAdd(a.ill, (unsigned long long)sill);
    }
}

// ---- render kernel: warp-private hash tables in shared memory --------------------------------------------------
// Detector images of imaging systems are PSF-like: the 7 M hits of the double-Gauss workload fall into 3e4 of
// the 4.5 M pixels, 2e5 of them into the hottest one, and fp64 atomics on a handful of L2 addresses serialise.
// After the warp-level aggregation of otb_bin.cuh (one leader lane per distinct pixel) every WARP accumulates
// its hits in its own shared-memory hash table (pixel index -> X, Y, Z, W sums and count; open addressing,
// 8 probes).  A table is only ever touched by the lanes of its warp and two leaders never hold the same pixel,
// so the sums are plain read-modify-writes: no fp64 atomics (shared memory has none natively; a CAS loop per
// channel was the bottleneck of a block-shared table).  Only claiming an empty slot is an atomic (two leaders of
// the warp may race for it).  At the end every occupied slot is flushed with one global atomic per channel; hits
// that find no slot (spread images overflow the table) go to global memory directly.
#define OTB_RH_WARPS 8
#define OTB_RH_SLOTS 256                       // per warp
#define OTB_RH_PROBES 8
#define OTB_RENDER_SMEM (OTB_RH_WARPS*OTB_RH_SLOTS*(4*sizeof(double) + 2*sizeof(int)))

__global__ void __launch_bounds__(32*OTB_RH_WARPS) render_kernel(BinGrid g, const double* __restrict__ obs, int64_t M,
                                                     const double* __restrict__ x, const double* __restrict__ y,
                                                     const float* __restrict__ w, const float* __restrict__ wl,
                                                     double* __restrict__ img, int* __restrict__ cnt)
{
    extern __shared__ double rh_smem[];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* val = rh_smem + (size_t)warp*4*OTB_RH_SLOTS;                                  // [4][SLOTS] of this warp
    int* key = (int*)(rh_smem + (size_t)OTB_RH_WARPS*4*OTB_RH_SLOTS) + warp*2*OTB_RH_SLOTS; // [SLOTS], -1 = empty
    int* num = key + OTB_RH_SLOTS;                                                        // [SLOTS]
    for (int i = lane; i < OTB_RH_SLOTS; i += 32) {
        key[i] = -1;
        num[i] = 0;
        val[i] = val[i + OTB_RH_SLOTS] = val[i + 2*OTB_RH_SLOTS] = val[i + 3*OTB_RH_SLOTS] = 0.0;
    }
    __syncwarp();

    // contiguous chunk per block (rays of one source are contiguous: fewer distinct pixels per table);
    // uniform trip count per warp: every lane takes part in the warp-level aggregation
    const int64_t chunk = ((M + gridDim.x - 1)/gridDim.x + blockDim.x - 1)/blockDim.x*blockDim.x;
    const int64_t lo = (int64_t)blockIdx.x*chunk, hi = (lo + chunk < M) ? lo + chunk : M;
    for (int64_t base = lo; base < hi; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const bool in = i < hi;
        const float wi = in ? w[i] : 0.0f;
        bool ok = in && (wi > 0.0f);            // only rays with a valid hit reach RenderImage.render
        double ox = 0.0, oy = 0.0, oz = 0.0, X = 0.0, Y = 0.0;
        if (ok) {
            X = x[i];
            Y = y[i];
            observer_xyz(obs, (double)wl[i], ox, oy, oz);
        }
        double v0, v1, v2, v3;
        int n, pix;
        ok = aggregate_xyz_warp(g, ok, X, Y, wi, ox, oy, oz, lane, v0, v1, v2, v3, n, pix);
        if (ok) {
            unsigned h = ((unsigned)pix*2654435761u) >> (32 - 8);           // Fibonacci hash, 256 slots
            int slot = -1;
#pragma unroll 1
            for (int t = 0; t < OTB_RH_PROBES; ++t) {
                int k = ((volatile int*)key)[h];
                if (k == -1) k = atomicCAS(&key[h], -1, pix), k = (k == -1) ? pix : k;
                if (k == pix) { slot = (int)h; break; }
                h = (h + 1) & (OTB_RH_SLOTS - 1);
            }
            if (slot >= 0) {
                val[slot] += v0;
                val[slot + OTB_RH_SLOTS] += v1;
                val[slot + 2*OTB_RH_SLOTS] += v2;
                val[slot + 3*OTB_RH_SLOTS] += v3;
                num[slot] += n;
            } else {
                double* q = img + 4*(int64_t)pix;
                atomicAdd(q + 0, v0);
                atomicAdd(q + 1, v1);
                atomicAdd(q + 2, v2);
                atomicAdd(q + 3, v3);
                if (cnt) atomicAdd(cnt + pix, n);
            }
        }
        __syncwarp();        // the table updates of this round are visible to the whole warp before the next one
    }
    for (int i = lane; i < OTB_RH_SLOTS; i += 32) {
        const int pix = key[i];
        if (pix >= 0) {
            double* q = img + 4*(int64_t)pix;
            atomicAdd(q + 0, val[i]);
            atomicAdd(q + 1, val[i + OTB_RH_SLOTS]);
            atomicAdd(q + 2, val[i + 2*OTB_RH_SLOTS]);
            atomicAdd(q + 3, val[i + 3*OTB_RH_SLOTS]);
            if (cnt) atomicAdd(cnt + pix, num[i]);
        }
    }
}

// per-device state (one slot per CUDA device: a process may drive several devices from several threads)
#define OTB_MAX_DEVICES 64
static double* g_obs_d[OTB_MAX_DEVICES] = {nullptr};
static bool g_render_smem_set[OTB_MAX_DEVICES] = {false};
static std::mutex g_dev_mutex;
int otb_sm_count();

int otb_observer_table(const double** out)
{
    int dev = 0;
    OTB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= OTB_MAX_DEVICES) { otb_set_error("device index out of range"); return OTB_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (!g_obs_d[dev]) {
        OTB_CUDA(cudaMalloc(&g_obs_d[dev], sizeof(OTB_OBSERVERS)));
        OTB_CUDA(cudaMemcpy(g_obs_d[dev], OTB_OBSERVERS, sizeof(OTB_OBSERVERS), cudaMemcpyHostToDevice));
    }
    *out = g_obs_d[dev];
    return OTB_OK;
}

BinGrid otb_make_grid(const double extent[4], int Nx, int Ny)
{
    BinGrid g;
    g.e0 = extent[0]; g.e1 = extent[1]; g.e2 = extent[2]; g.e3 = extent[3];
    g.Nx = Nx; g.Ny = Ny;
    g.fx = (double)Nx/(extent[1] - extent[0]);
    g.fy = (double)Ny/(extent[3] - extent[2]);
    return g;
}

extern "C" {

int otb_detector_hits(const OtbRayStore* store, int64_t ray_begin, int64_t ray_end, const OtbDetector* det_h,
                      double* hx_d, double* hy_d, float* hw_d, double* range_d, int64_t* ill_d, int32_t* status_d, void* stream)
{
    if (!store || !det_h || !hx_d || !hy_d || !hw_d || !range_d || !ill_d || !status_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (ray_begin < 0 || ray_end > store->N || ray_begin > ray_end) { otb_set_error("invalid ray range"); return OTB_ERR_INVALID_ARG; }
    const int k = det_h->surface.kind;
    if (k == OTB_SURF_FUNC || k == OTB_SURF_DATA || k == OTB_SURF_ASPHERE) {
        otb_set_error("Function/Data surfaces are not supported as detector surfaces (detector.py:37-41)");
        return OTB_ERR_UNSUPPORTED;
    }
    const int64_t n = ray_end - ray_begin;
    if (n == 0) return OTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DetArgs a;
    a.st = *store;
    a.surf = otb_ksurface(det_h->surface);
    a.projection = det_h->projection;
    a.has_extent = det_h->has_extent;
    for (int i = 0; i < 4; ++i) a.extent[i] = det_h->extent[i];
    a.begin = ray_begin;
    a.end = ray_end;
    a.hx = hx_d; a.hy = hy_d; a.hw = hw_d;
    a.range = range_d;
    a.ill = (unsigned long long*)ill_d;
    a.status = status_d;
    detector_hits_kernel<<<(unsigned)((n + OTB_DET_THREADS - 1)/OTB_DET_THREADS), OTB_DET_THREADS, 0, st>>>(a);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

int otb_render_xyzw(const double* x_d, const double* y_d, const float* w_d, const float* wl_d, int64_t M,
                    const double extent[4], int32_t Nx, int32_t Ny, double* img_d, int32_t* cnt_d, void* stream)
{
    if (!extent || !img_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (Nx <= 0 || Ny <= 0 || !(extent[1] > extent[0]) || !(extent[3] > extent[2])) { otb_set_error("invalid image grid"); return OTB_ERR_INVALID_ARG; }
    if (M <= 0) return OTB_OK;
    if (!x_d || !y_d || !w_d || !wl_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    const double* obs;
    if (int rc = otb_observer_table(&obs)) return rc;
    BinGrid g = otb_make_grid(extent, Nx, Ny);
    int dev = 0;
    OTB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= OTB_MAX_DEVICES) { otb_set_error("device index out of range"); return OTB_ERR_INVALID_ARG; }
    if (!g_render_smem_set[dev]) {
        OTB_CUDA(cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OTB_RENDER_SMEM));
        g_render_smem_set[dev] = true;
    }
    const int blocks = otb_one_wave_grid(render_kernel, 256, OTB_RENDER_SMEM, otb_sm_count(), (M + 255)/256);
    render_kernel<<<blocks, 256, OTB_RENDER_SMEM, (cudaStream_t)stream>>>(g, obs, M, x_d, y_d, w_d, wl_d, img_d, cnt_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

}  // extern "C"
