// otb_render.cu — fused render mode: trace + detector test + XYZW binning with NO per-surface storage — the
// per-chunk body of Raytracer.iterative_render (raytracer.py:1235-1264).
//
// The detector logic of Raytracer._hit_detector (raytracer.py:929-991) is evaluated online while the ray
// segment i -> i+1 is still in registers: the walk starts at the section before the first stored point with
// z >= z_min of the detector, a hit is valid unless it lies behind the next stored point (+C_EPS), and rays
// whose points lie all behind or all before the detector's z-extent are ignored.  Segment directions are
// recomputed from position differences exactly like RayStorage.rays_by_mask (ray_storage.py:273-285), so the
// binned hits are identical to those of the store path.
#include <string.h>
#include "otb_step.cuh"
#include "otb_gen.cuh"
#include "otb_bin.cuh"

#define OTB_RENDER_THREADS 128
#define OTB_MAX_DET 8

struct RenderDet {
    KSurface surf;
    int projection, has_extent;
    double extent[4];
    BinGrid grid;
    double* img;
    int* cnt;
    double* range;      // [4] hit range accumulation (mode 1)
};

struct RenderArgs {
    KScene sc;
    OtbRays in;
    RenderDet dets[OTB_MAX_DET];      // by value (constant bank)
    int n_det;
    int mode;            // 0: bin into the images, 1: only accumulate the hit ranges (auto extent)
    const double* obs;
    unsigned long long* msgs;
    int* status;
    int nt;
    int64_t k_begin, k_end;     // ray range of this launch
    GenBlock G;                 // G.nsrc > 0: rays are generated here (otb_gen.cuh)
};
static_assert(sizeof(RenderArgs) <= 32764, "kernel parameter block exceeds the 32 KB limit");

// Per-ray detector bookkeeping: two bit masks (bit d = detector d) instead of per-detector records — the walk of a
// detector starts once (first stored point at or behind its z_min) and finishes once (hit accepted or rejected), and a
// finished hit is binned at once while the segment is in registers, so nothing per detector has to survive the step.
// Rays whose FIRST point already lies behind a detector's z-extent are ignored for it like in the reference
// (raytracer.py:929-938, "no_start": with z non-decreasing along a ray the first point decides; a ray with decreasing
// z sets OTB_STATUS_Z_DECREASE and the caller repeats the chunk through the stored-section path).
struct DetMasks {
    unsigned started, finished, skipped;
};

// one sequential step of the fused mode: trace, book messages, then the detector walk of _hit_detector
// (raytracer.py:881-1051) online on section i = (p_i -> r.p) while it is still in registers
template <bool POL, int CAPS>
__device__ __forceinline__ void render_step(const KScene& sc, const RenderArgs& a, const int i, RayState& r, DetMasks& dm,
                                            unsigned long long* srng, bool& z_decrease,
                                            int* smsgs, const bool valid, const int64_t ray)
{
    const int64_t N = a.in.N;
    const int NDET = a.n_det;
    const OtbStep& st = sc.steps[i];
    double za = 0.0, zb = 0.0;
    if (CAPS == OTB_CAPS_FULL && st.hurb && valid) {
        if (a.in.hurb_z_d) {
            za = a.in.hurb_z_d[((int64_t)st.hurb_slot*2 + 0)*N + ray];
            zb = a.in.hurb_z_d[((int64_t)st.hurb_slot*2 + 1)*N + ray];
        } else {
            Philox4 rnd = philox4x32_10((uint64_t)(a.in.ray_offset + ray), 0x48555242u, (uint32_t)st.hurb_slot, a.in.seed);
            normal2(rnd, za, zb);
        }
    }
    const V3 p_i = r.p;
    const float w_i = r.w;
    StepFlags fl;
    trace_step<POL, CAPS>(sc, a.sc.aux, st, r, fl, za, zb, a.status);
    book_step(smsgs, a.nt, i, valid, fl);
    z_decrease = z_decrease | (r.p.z < p_i.z);

    // which detectors does this segment concern?  (warp-uniform decision: the binning below is a warp collective)
    unsigned active = 0;
    for (int d = 0; d < NDET; ++d) {
        const bool bmin = r.p.z >= a.dets[d].surf.z_min;
        if (bmin) dm.started |= 1u << d;                 // section before the first point behind z_min
    }
    active = valid ? (dm.started & ~dm.finished & ~dm.skipped) : 0u;
    unsigned any = active;
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, k);
    if (!any) return;
    const V3 sd = unit3(v3(r.p.x - p_i.x, r.p.y - p_i.y, r.p.z - p_i.z));      // shared by all detectors
    double ox = 0.0, oy = 0.0, oz = 0.0;
    bool have_obs = false;
    for (int d = 0; d < NDET; ++d) {
        if (!((any >> d) & 1u)) continue;
        const RenderDet& rd = a.dets[d];
        bool ok = false;
        double X = 0.0, Y = 0.0;
        if ((active >> d) & 1u) {
            HitResult h = surf_find_hit<(CAPS == OTB_CAPS_LENS ? OTB_CAPS_LENS : OTB_CAPS_DET)>(rd.surf, nullptr, p_i, sd, a.status);
            if (!(h.p.z > r.p.z + OTB_C_EPS)) {           // else: hit behind the next stored point, next section
                dm.finished |= 1u << d;
                if (h.hit && w_i > 0.0f) {
                    X = h.p.x;
                    Y = h.p.y;
                    sphere_project(rd.surf, rd.projection, X, Y, h.p.z);
                    ok = true;
                    if (rd.has_extent) {
                        const double* e = rd.extent;
                        ok = (e[0] <= X) && (X <= e[1]) && (e[2] <= Y) && (Y <= e[3]);
                    }
                }
            }
        }
        if (a.mode == 0) {
            if (__any_sync(0xffffffffu, ok)) {
                if (!have_obs) {
                    observer_xyz(a.obs, (double)r.wl, ox, oy, oz);      // one observer lookup per ray and step at most
                    have_obs = true;
                }
                accumulate_xyz_warp(rd.grid, ok, X, Y, w_i, ox, oy, oz, rd.img, rd.cnt);
            }
        } else if (__any_sync(0xffffffffu, ok)) {
            // range mode (auto extent of the first chunk): warp reduction, order-preserving keys in shared memory
            double mnx = ok ? X : INFINITY, mxx = ok ? X : -INFINITY, mny = ok ? Y : INFINITY, mxy = ok ? Y : -INFINITY;
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) {
                mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, k));
                mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, k));
                mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, k));
                mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, k));
            }
            if ((threadIdx.x & 31) == 0) {
                atomicMin(&srng[4*d + 0], dkey(mnx));
                atomicMax(&srng[4*d + 1], dkey(mxx));
                atomicMin(&srng[4*d + 2], dkey(mny));
                atomicMax(&srng[4*d + 3], dkey(mxy));
            }
        }
    }
}

// GEN: rays are drawn in the kernel (a template parameter: no generator code in the kernels of pre-generated bundles)
template <bool POL, int CAPS, bool GEN>
__global__ void __launch_bounds__(OTB_RENDER_THREADS, OTB_MINBLOCKS(CAPS))
trace_render_kernel(const __grid_constant__ RenderArgs a)
{
    extern __shared__ int smsgs[];
#if OTB_SPEC
    const KScene& sc = K_SPEC;
#else
    const KScene& sc = a.sc;
#endif
    const double* __restrict__ aux = a.sc.aux;
    const int nt = a.nt;
    const int64_t N = a.in.N;
    const int NDET = a.n_det;
    __shared__ unsigned long long srng[4*OTB_MAX_DET];      // range mode: hit ranges of the block (min, max, min, max keys)
    for (int i = threadIdx.x; i < OTB_NMSG*nt; i += blockDim.x) smsgs[i] = 0;
    if (threadIdx.x < 4*OTB_MAX_DET) srng[threadIdx.x] = (threadIdx.x & 1) ? 0ull : ~0ull;
    __syncthreads();

    for (int64_t base = a.k_begin + (int64_t)blockIdx.x*blockDim.x; base < a.k_end; base += (int64_t)gridDim.x*blockDim.x) {
        const int64_t ray = base + threadIdx.x;
        const bool valid = ray < a.k_end;
        RayState r;
        if (GEN) {
            GenRay gr;
            generate_ray(a.G, valid ? ray : a.k_begin, gr);
            if (valid && gr.neg_dir) atomicOr(a.status, OTB_STATUS_NEG_DIR);
            r.p = gr.p;
            r.s = gr.s;
            r.w = valid ? gr.w : 0.0f;
            r.wl = gr.wl;
            r.pol[0] = POL ? gr.pol[0] : 0.0f;
            r.pol[1] = POL ? gr.pol[1] : 0.0f;
            r.pol[2] = POL ? gr.pol[2] : 0.0f;
        } else if (valid) {
            r.p = v3(a.in.p0_d[ray], a.in.p0_d[ray + N], a.in.p0_d[ray + 2*N]);
            r.s = v3(a.in.s0_d[ray], a.in.s0_d[ray + N], a.in.s0_d[ray + 2*N]);
            r.w = a.in.w0_d[ray];
            r.wl = a.in.wl_d[ray];
            if (POL) {
                r.pol[0] = a.in.pol0_d[ray];
                r.pol[1] = a.in.pol0_d[ray + N];
                r.pol[2] = a.in.pol0_d[ray + 2*N];
            }
        } else {
            r.p = v3(0, 0, 0);
            r.s = v3(0, 0, 1);
            r.w = 0.0f;
            r.wl = 550.0f;
            r.pol[0] = r.pol[1] = r.pol[2] = 0.0f;
        }
        r.n = medium_n(sc.media[sc.medium0], aux, (double)r.wl);
        if (valid && r.n < 1.0) atomicOr(a.status, OTB_STATUS_NBELOW1);

        DetMasks dm;
        dm.started = dm.finished = dm.skipped = 0u;
        for (int d = 0; d < NDET; ++d) {
            const KSurface& D = a.dets[d].surf;
            const bool bmin = r.p.z >= D.z_min, bmax = r.p.z >= D.z_max;
            if (bmin) dm.started |= 1u << d;      // first stored point already at/behind z_min: the walk starts at section 0
            if (bmin && bmax) dm.skipped |= 1u << d;          // ray starts behind the detector: ignored (no_start)
        }
        bool z_decrease = false;

#if OTB_SPEC
#define OTB_CALL_RENDER_STEP(i) render_step<POL, CAPS>(sc, a, i, r, dm, srng, z_decrease, smsgs, valid, ray);
        OTB_SPEC_FOREACH_STEP(OTB_CALL_RENDER_STEP)      // straight-line code, see otb_trace.cu
#else
        for (int i = 0; i < sc.n_steps; ++i) render_step<POL, CAPS>(sc, a, i, r, dm, srng, z_decrease, smsgs, valid, ray);
#endif
        // rays still walking at the last stored point have no further section: no hit (raytracer.py:970-978)
        if (valid && z_decrease) atomicOr(a.status, OTB_STATUS_Z_DECREASE);
    }

    __syncthreads();
    if (a.mode != 0 && threadIdx.x < NDET) {
        const int d = threadIdx.x;
        if (srng[4*d] != ~0ull) {
            atomic_min_double(&a.dets[d].range[0], dkey_inv(srng[4*d + 0]));
            atomic_max_double(&a.dets[d].range[1], dkey_inv(srng[4*d + 1]));
            atomic_min_double(&a.dets[d].range[2], dkey_inv(srng[4*d + 2]));
            atomic_max_double(&a.dets[d].range[3], dkey_inv(srng[4*d + 3]));
        }
    }

    __syncthreads();
    for (int i = threadIdx.x; i < OTB_NMSG*nt; i += blockDim.x)
        if (smsgs[i]) atomicAdd(&a.msgs[i], (unsigned long long)smsgs[i]);
}

#define OTB_LAUNCH_RENDER_G(POL, CAPS, GEN) do { \
        int blocks = otb_one_wave_grid(trace_render_kernel<POL, CAPS, GEN>, OTB_RENDER_THREADS, smem, otb_sm_count(), blocks_needed); \
        trace_render_kernel<POL, CAPS, GEN><<<blocks, OTB_RENDER_THREADS, smem, st>>>(a); } while (0)
#define OTB_LAUNCH_RENDER(POL, CAPS) do { if (a.G.nsrc > 0) OTB_LAUNCH_RENDER_G(POL, CAPS, true); \
        else OTB_LAUNCH_RENDER_G(POL, CAPS, false); } while (0)

int otb_observer_table(const double** out);
int otb_check_sources(const OtbSource* sources_h, int n_sources, int64_t N);
void otb_fill_genblock(GenBlock* G, const OtbSource* sources_h, int g0, int nsrc, const double* gen_aux_d, uint64_t seed,
                       int64_t ray_offset, int no_pol);
BinGrid otb_make_grid(const double extent[4], int Nx, int Ny);
int otb_sm_count();

extern "C" int otb_trace_render(const OtbScene* scene, const OtbRays* rays, int n_det, const OtbDetector* dets_h,
                                const double* extents_h, const int32_t* Nx_h, const int32_t* Ny_h,
                                double* const* img_d, int32_t* const* cnt_d, double* range_d,
                                int64_t* msgs_d, int32_t* status_d, void* stream)
{
    if (!scene || !rays || !dets_h || !msgs_d || !status_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (n_det < 1 || n_det > OTB_MAX_DET) { otb_set_error("between 1 and %d detectors per fused launch", OTB_MAX_DET); return OTB_ERR_INVALID_ARG; }
    const int mode = (img_d == nullptr) ? 1 : 0;
    if (mode == 1 && !range_d) { otb_set_error("range_d required when no images are given"); return OTB_ERR_INVALID_ARG; }
    if (mode == 0 && (!extents_h || !Nx_h || !Ny_h)) { otb_set_error("extents and grid sizes required"); return OTB_ERR_INVALID_ARG; }
    if (rays->N <= 0) return OTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    RenderArgs a;
    memset(a.dets, 0, sizeof(a.dets));
    RenderDet* hd = a.dets;
    for (int d = 0; d < n_det; ++d) {
        const int k = dets_h[d].surface.kind;
        if (k == OTB_SURF_FUNC || k == OTB_SURF_DATA || k == OTB_SURF_ASPHERE) {
            otb_set_error("Function/Data surfaces are not supported as detector surfaces (detector.py:37-41)");
            return OTB_ERR_UNSUPPORTED;
        }
        hd[d].surf = otb_ksurface(dets_h[d].surface);
        hd[d].projection = dets_h[d].projection;
        hd[d].has_extent = dets_h[d].has_extent;
        for (int q = 0; q < 4; ++q) hd[d].extent[q] = dets_h[d].extent[q];
        if (mode == 0) {
            const double* e = extents_h + 4*d;
            if (Nx_h[d] <= 0 || Ny_h[d] <= 0 || !(e[1] > e[0]) || !(e[3] > e[2]) || !img_d[d]) {
                otb_set_error("invalid image grid for detector %d", d);
                return OTB_ERR_INVALID_ARG;
            }
            hd[d].grid = otb_make_grid(e, Nx_h[d], Ny_h[d]);
            hd[d].img = img_d[d];
            hd[d].cnt = cnt_d ? cnt_d[d] : nullptr;
        } else {
            hd[d].range = range_d + 4*d;
        }
    }
    a.sc = scene->k;
    a.in = *rays;
    a.n_det = n_det;
    a.mode = mode;
    if (int rc = otb_observer_table(&a.obs)) return rc;
    a.msgs = (unsigned long long*)msgs_d;
    a.status = status_d;
    a.nt = scene->nt;
    const int64_t N = rays->N;
    size_t smem = sizeof(int)*OTB_NMSG*scene->nt;
    bool lean = scene->caps == OTB_CAPS_LENS;
    for (int d = 0; d < n_det; ++d) if (dets_h[d].surface.kind == OTB_SURF_TILTED) lean = false;
    // ray ranges: one launch for an injected bundle, one per group of <= OTB_GEN_MAXSRC sources when generating
    const OtbGenerator* gen = rays->gen_h;
    if (gen) { if (int rc = otb_check_sources(gen->sources_h, gen->n_sources, N)) return rc; }
    else if (!rays->p0_d || !rays->s0_d || !rays->w0_d || !rays->wl_d || (!scene->k.no_pol && !rays->pol0_d)) {
        otb_set_error("missing ray array");
        return OTB_ERR_INVALID_ARG;
    }
    const int n_groups = gen ? (gen->n_sources + OTB_GEN_MAXSRC - 1)/OTB_GEN_MAXSRC : 1;
    for (int grp = 0; grp < n_groups; ++grp) {
    if (gen) {
        const int g0 = grp*OTB_GEN_MAXSRC;
        const int nsrc = (gen->n_sources - g0 < OTB_GEN_MAXSRC) ? gen->n_sources - g0 : OTB_GEN_MAXSRC;
        otb_fill_genblock(&a.G, gen->sources_h, g0, nsrc, gen->gen_aux_d, rays->seed, rays->ray_offset, scene->k.no_pol);
        a.k_begin = gen->sources_h[g0].ray_start;
        a.k_end = gen->sources_h[g0 + nsrc - 1].ray_start + gen->sources_h[g0 + nsrc - 1].n_rays;
        if (a.k_end <= a.k_begin) continue;
    } else {
        a.G.nsrc = 0;
        a.k_begin = 0;
        a.k_end = N;
    }
    const int64_t blocks_needed = (a.k_end - a.k_begin + OTB_RENDER_THREADS - 1)/OTB_RENDER_THREADS;
#if OTB_SPEC
    if (!otb_scene_equal(scene->k, K_SPEC_HOST)) {
        otb_set_error("this engine build is specialised for a different scene");
        return OTB_ERR_INVALID_ARG;
    }
    if (lean || OTB_SPEC_CAPS == OTB_CAPS_FULL) OTB_LAUNCH_RENDER((OTB_SPEC_POL != 0), OTB_SPEC_CAPS);
    else OTB_LAUNCH_RENDER((OTB_SPEC_POL != 0), OTB_CAPS_FULL);
#else
    if (scene->k.no_pol) {
        if (lean) OTB_LAUNCH_RENDER(false, OTB_CAPS_LENS);
        else OTB_LAUNCH_RENDER(false, OTB_CAPS_FULL);
    } else {
        if (lean) OTB_LAUNCH_RENDER(true, OTB_CAPS_LENS);
        else OTB_LAUNCH_RENDER(true, OTB_CAPS_FULL);
    }
#endif
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return otb_cuda_fail(e, "trace_render_kernel launch");
    }
    return OTB_OK;
}
