// otb_trace.cu — store-mode trace kernel: the whole surface loop of Raytracer.trace (raytracer.py:297-397)
// in ONE launch.  Thread = ray, ray state in registers, every section written once as coalesced SoA planes
// in the byte layout of RayStorage (ray_storage.py:80-90).  No tensor cores: the path is not a contraction.
//
// The scene (steps, surfaces, media, filters) is a __grid_constant__ kernel parameter: warp-uniform values come
// from the constant bank, not from global memory.
// HBM traffic per ray: write nt*48 + 28 B (pol) or nt*36 + 28 B (no_pol); read 68 B only for INJECTED bundles —
// with OtbRays.gen_h the kernel draws its rays itself (generate_ray, otb_gen.cuh) and the bundle never exists in HBM.
#include "otb_step.cuh"
#include "otb_gen.cuh"

#ifndef OTB_TRACE_THREADS
#define OTB_TRACE_THREADS 128
#endif
// Scenes with numeric surfaces (CAPS_FULL) run ONE block of 512 threads per SM (16 warps at 128 registers; measured on
// cosine_surfaces / zoo_numeric: 256 threads 11.2 / 120 ms, 384: 8.7 / 106, 448: 9.2 / 106, 512: 7.6 / 89, 640: 9.5 / 89,
// 768: 10.5 / 91 ms per 10 M rays — the kernel is latency-bound, more warps help until the spills take over): the
// block then owns the SM's shared memory and stages the scene's aux tables in it — spline
// knots / coefficients of DataSurfaces, asphere polynomials, tabulated media and filters — with ONE bulk asynchronous
// copy (TMA, cp.async.bulk + mbarrier) when they fit; every table lookup of the Illinois iteration and of the normal
// (a 5 x 5 coefficient patch plus knots per spline evaluation, ~13 evaluations per hit) then reads shared memory
// instead of L1/L2 (data_surface_2d.py:104, 130-153).  Larger tables stay in global memory (L2 resident).
#ifndef OTB_TRACE_THREADS_FULL
#define OTB_TRACE_THREADS_FULL 512      // scenes with asphere / function / data surfaces
#endif
#ifndef OTB_TRACE_THREADS_FULL_FLAT
#define OTB_TRACE_THREADS_FULL_FLAT 384 // CAPS_FULL for other reasons (HURB apertures, tilted planes, tabulated media): 168 registers
#endif
#define OTB_AUX_SMEM_MAX (200*1024)
#define OTB_BLOCKS_OF(CAPS) ((CAPS) == OTB_CAPS_FULL ? 1 : OTB_MINBLOCKS(CAPS))

// one bulk asynchronous copy global -> shared (TMA engine), completion on an mbarrier; bytes: multiple of 16
__device__ __forceinline__ void tma_stage(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* mbar)
{
    const unsigned bar = (unsigned)__cvta_generic_to_shared(mbar);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(dst_smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        // chunks of at most 64 KB
        for (unsigned off = 0; off < bytes; off += 65536u) {
            const unsigned n = (bytes - off < 65536u) ? bytes - off : 65536u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst + off), "l"((const char*)src_gmem + off), "r"(n), "r"(bar) : "memory");
        }
    }
    // every thread waits for phase 0 of the barrier: the data is visible to it afterwards
    asm volatile("{\n\t.reg .pred p;\n\tOTB_TMA_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra OTB_TMA_WAIT;\n\t}"
                 :: "r"(bar) : "memory");
}
#ifndef OTB_TRACE_MINBLOCKS
#define OTB_TRACE_MINBLOCKS 3      // resident blocks per SM the register allocation aims at (see profiles/)
#endif

struct TraceArgs {
    KScene sc;
    OtbRays in;
    OtbRayStore out;
    unsigned long long* msgs;   // [OTB_NMSG * nt]
    int* status;
    int aux_smem_bytes, pad0;   // > 0: stage that many bytes of the aux tables in shared memory (16-byte multiple)
    int64_t k_begin, k_end;     // ray range of this launch (more than OTB_GEN_MAXSRC sources: one launch per group)
    GenBlock G;                 // G.nsrc > 0: rays are generated here
};
static_assert(sizeof(TraceArgs) <= 32764, "kernel parameter block exceeds the 32 KB limit");

#ifndef OTB_STEP_UNROLL
#define OTB_STEP_UNROLL 1
#endif
#define OTB_STR(x) #x
#define OTB_UNROLL_N(n) _Pragma(OTB_STR(unroll n))

struct StoreCursor {
    double* pp;     // p plane x of the current section (y, z at +N*nt, +2*N*nt)
    float* pw;
    double* pn;
    float* ppol;
    bool z_decrease;   // some section of this ray ended at a smaller z than it started
};

// one sequential step: trace, book messages, store section i+1
template <bool POL, int CAPS>
__device__ __forceinline__ void store_step(const KScene& sc, const TraceArgs& a, const double* aux, const int i, RayState& r,
                                           StoreCursor& c, int* smsgs, const bool valid, const int64_t rr, const bool func_spec)
{
    const int64_t N = a.out.N;
    const int nt = a.out.nt;
    const int64_t Nnt = N*(int64_t)nt;
    const OtbStep& st = sc.steps[i];
    double za = 0.0, zb = 0.0;
    if (CAPS == OTB_CAPS_FULL && st.hurb) {
        if (a.in.hurb_z_d) {
            za = a.in.hurb_z_d[((int64_t)st.hurb_slot*2 + 0)*N + rr];
            zb = a.in.hurb_z_d[((int64_t)st.hurb_slot*2 + 1)*N + rr];
        } else {
            Philox4 rnd = philox4x32_10((uint64_t)(a.in.ray_offset + rr), 0x48555242u, (uint32_t)st.hurb_slot, a.in.seed);
            normal2(rnd, za, zb);
        }
    }
    StepFlags fl;
    const double z_prev = r.p.z;
    trace_step<POL, CAPS>(sc, aux, st, r, fl, za, zb, a.status, func_spec);
    c.z_decrease = c.z_decrease | (r.p.z < z_prev);
    book_step(smsgs, nt, i, valid, fl);

    c.pp += N;
    c.pw += N;
    c.pn += N;
    if (valid) {
        __stcs(c.pp, r.p.x);
        __stcs(c.pp + Nnt, r.p.y);
        __stcs(c.pp + 2*Nnt, r.p.z);
        __stcs(c.pw, r.w);
        __stcs(c.pn, r.n);
    }
    if (POL) {
        c.ppol += N;
        if (valid) {
            __stcs(c.ppol, r.pol[0]);
            __stcs(c.ppol + Nnt, r.pol[1]);
            __stcs(c.ppol + 2*Nnt, r.pol[2]);
        }
    }
}

// GEN: the rays are drawn inside the kernel (fused RaySource.create_rays).  A template parameter, not a run-time
// branch: the generator is ~5000 instructions, and the kernels of injected / pre-generated bundles (the default) are
// sensitive to their code footprint (measured on the fused render kernel, where only this changed: 6-detector bin
// pass 1.87 -> 1.58 ms per 10 M rays; cosine_surfaces 12.2 -> 9.5 ms together with the out-of-line height functions).
// THREADS: block size the register allocation is made for (see above; measured on hurb_square / hurb_pinhole:
// 384 threads 2.27 / 1.78 ms, 512 threads 2.43 / 2.04 ms per 10 M rays — the opposite of the numeric-surface scenes).
template <bool POL, int CAPS, bool GEN, int THREADS>
__global__ void __launch_bounds__(THREADS, OTB_BLOCKS_OF(CAPS))
trace_store_kernel(const __grid_constant__ TraceArgs a)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    // layout: [aux tables (aux_smem_bytes, 16-byte aligned)] [mbarrier (8)] [pad (8)] [message counters]
    int* smsgs = (int*)(dyn_smem + a.aux_smem_bytes + 16);      // [OTB_NMSG * nt]
#if OTB_SPEC
    const KScene& sc = K_SPEC;
#else
    const KScene& sc = a.sc;
#endif
    const double* aux = a.sc.aux;
    const int nt = a.out.nt;
    const int64_t N = a.out.N;
    for (int i = threadIdx.x; i < OTB_NMSG*nt; i += blockDim.x) smsgs[i] = 0;
    if (CAPS == OTB_CAPS_FULL && a.aux_smem_bytes > 0) {
        tma_stage(dyn_smem, a.sc.aux, (unsigned)a.aux_smem_bytes, (unsigned long long*)(dyn_smem + a.aux_smem_bytes));
        aux = (const double*)dyn_smem;            // generic pointer into shared memory: the table code is unchanged
    }
    __syncthreads();

    const int64_t Nnt = N*(int64_t)nt;
    // all numeric surfaces of the scene are function surfaces: they take their kind-specialised step (trace_step)
    bool func_spec = (CAPS == OTB_CAPS_FULL);
    if (CAPS == OTB_CAPS_FULL) {
        for (int i = 0; i < sc.n_steps; ++i) {
            const KSurface& S = sc.surf[sc.steps[i].surface];
            if (!(S.flags & OTB_SF_FLAT) && (S.kind == OTB_SURF_DATA || S.kind == OTB_SURF_ASPHERE)) func_spec = false;
        }
    }

    for (int64_t base = a.k_begin + (int64_t)blockIdx.x*blockDim.x; base < a.k_end; base += (int64_t)gridDim.x*blockDim.x) {
        const int64_t ray = base + threadIdx.x;
        const bool valid = ray < a.k_end;
        const int64_t rr = valid ? ray : a.k_begin;  // clamp so that every lane addresses valid memory
        RayState r;
        if (GEN) {
            // fused RaySource.create_rays: the ray is drawn here (Philox counter = global ray id) instead of being
            // written by a generator kernel and read back
            GenRay gr;
            generate_ray(a.G, rr, gr);
            if (valid && gr.neg_dir) atomicOr(a.status, OTB_STATUS_NEG_DIR);
            r.p = gr.p;
            r.s = gr.s;
            r.w = valid ? gr.w : 0.0f;
            r.wl = gr.wl;
            r.pol[0] = POL ? gr.pol[0] : 0.0f;
            r.pol[1] = POL ? gr.pol[1] : 0.0f;
            r.pol[2] = POL ? gr.pol[2] : 0.0f;
        } else {
            r.p = v3(a.in.p0_d[rr], a.in.p0_d[rr + N], a.in.p0_d[rr + 2*N]);
            r.s = v3(a.in.s0_d[rr], a.in.s0_d[rr + N], a.in.s0_d[rr + 2*N]);
            r.w = valid ? a.in.w0_d[rr] : 0.0f;
            r.wl = a.in.wl_d[rr];
            if (POL) {
                r.pol[0] = a.in.pol0_d[rr];
                r.pol[1] = a.in.pol0_d[rr + N];
                r.pol[2] = a.in.pol0_d[rr + 2*N];
            } else {
                r.pol[0] = r.pol[1] = r.pol[2] = 0.0f;
            }
        }
        r.n = medium_n(sc.media[sc.medium0], aux, (double)r.wl);
        if (valid && r.n < 1.0) atomicOr(a.status, OTB_STATUS_NBELOW1);

        // running plane pointers: one add per plane and section instead of 64-bit index arithmetic
        StoreCursor cur;
        cur.z_decrease = false;
        cur.pp = a.out.p_d + rr;
        cur.pw = a.out.w_d + rr;
        cur.pn = a.out.n_d + rr;
        cur.ppol = POL ? a.out.pol_d + rr : nullptr;
        double* const pp = cur.pp;
        float* const pw = cur.pw;
        double* const pn = cur.pn;
        float* const ppol = cur.ppol;

        if (valid) {
            __stcs(pp, r.p.x);
            __stcs(pp + Nnt, r.p.y);
            __stcs(pp + 2*Nnt, r.p.z);
            __stcs(pw, r.w);
            __stcs(pn, r.n);
            __stcs(&a.out.wl_d[rr], r.wl);
            if (POL) {
                __stcs(ppol, r.pol[0]);
                __stcs(ppol + Nnt, r.pol[1]);
                __stcs(ppol + 2*Nnt, r.pol[2]);
            }
        }

#if OTB_SPEC
        // straight-line code: one inlined copy of store_step per literal step index, everything about the step
        // (role, surface kind and parameters, media) folds at compile time
#define OTB_CALL_STORE_STEP(i) store_step<POL, CAPS>(sc, a, aux, i, r, cur, smsgs, valid, rr, func_spec);
        OTB_SPEC_FOREACH_STEP(OTB_CALL_STORE_STEP)
#else
        OTB_UNROLL_N(OTB_STEP_UNROLL)
        for (int i = 0; i < sc.n_steps; ++i) store_step<POL, CAPS>(sc, a, aux, i, r, cur, smsgs, valid, rr, func_spec);
#endif
        if (valid && cur.z_decrease) atomicOr(a.status, OTB_STATUS_Z_DECREASE);    // practically never
        if (valid) {
            __stcs(&a.out.s_d[rr], r.s.x);
            __stcs(&a.out.s_d[rr + N], r.s.y);
            __stcs(&a.out.s_d[rr + 2*N], r.s.z);
        }
    }

    __syncthreads();
    for (int i = threadIdx.x; i < OTB_NMSG*nt; i += blockDim.x)
        if (smsgs[i]) atomicAdd(&a.msgs[i], (unsigned long long)smsgs[i]);
}

int otb_check_sources(const OtbSource* sources_h, int n_sources, int64_t N);
void otb_fill_genblock(GenBlock* G, const OtbSource* sources_h, int g0, int nsrc, const double* gen_aux_d, uint64_t seed,
                       int64_t ray_offset, int no_pol);

static int launch_trace_store_range(const OtbScene* scene, TraceArgs& a, cudaStream_t stream, int sm_count);

int otb_launch_trace_store(const OtbScene* scene, const OtbRays* rays, const OtbRayStore* out,
                           int64_t* msgs_d, int32_t* status_d, cudaStream_t stream, int sm_count)
{
    TraceArgs a;
    a.sc = scene->k;
    a.in = *rays;
    a.out = *out;
    a.msgs = (unsigned long long*)msgs_d;
    a.status = status_d;
    const int64_t N = out->N;
    if (N <= 0) return OTB_OK;
    if (!rays->gen_h) {
        a.G.nsrc = 0;
        a.k_begin = 0;
        a.k_end = N;
        return launch_trace_store_range(scene, a, stream, sm_count);
    }
    const OtbGenerator& gen = *rays->gen_h;
    if (int rc = otb_check_sources(gen.sources_h, gen.n_sources, N)) return rc;
    for (int g0 = 0; g0 < gen.n_sources; g0 += OTB_GEN_MAXSRC) {
        const int nsrc = (gen.n_sources - g0 < OTB_GEN_MAXSRC) ? gen.n_sources - g0 : OTB_GEN_MAXSRC;
        otb_fill_genblock(&a.G, gen.sources_h, g0, nsrc, gen.gen_aux_d, rays->seed, rays->ray_offset, scene->k.no_pol);
        a.k_begin = gen.sources_h[g0].ray_start;
        a.k_end = gen.sources_h[g0 + nsrc - 1].ray_start + gen.sources_h[g0 + nsrc - 1].n_rays;
        if (a.k_end <= a.k_begin) continue;
        if (int rc = launch_trace_store_range(scene, a, stream, sm_count)) return rc;
    }
    return OTB_OK;
}

template <bool POL, int CAPS, bool GEN, int THREADS>
static int launch_store_threads(const TraceArgs& a, cudaStream_t stream, int sm_count, size_t smem)
{
    const int64_t blocks_needed = (a.k_end - a.k_begin + THREADS - 1)/THREADS;
    if (smem > 48*1024) {
        cudaError_t ea = cudaFuncSetAttribute(trace_store_kernel<POL, CAPS, GEN, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ea != cudaSuccess) return otb_cuda_fail(ea, "cudaFuncSetAttribute(trace_store_kernel)");
    }
    const int blocks = otb_one_wave_grid(trace_store_kernel<POL, CAPS, GEN, THREADS>, THREADS, smem, sm_count, blocks_needed);
    trace_store_kernel<POL, CAPS, GEN, THREADS><<<blocks, THREADS, smem, stream>>>(a);
    return OTB_OK;
}

// block size per capability level; fused generation keeps one block size (fewer instantiations of the largest kernels)
template <bool POL, int CAPS, bool GEN>
static int launch_store_caps(const TraceArgs& a, cudaStream_t stream, int sm_count, size_t smem, bool numeric)
{
    if constexpr (CAPS != OTB_CAPS_FULL) return launch_store_threads<POL, CAPS, GEN, OTB_TRACE_THREADS>(a, stream, sm_count, smem);
    else if constexpr (GEN) return launch_store_threads<POL, CAPS, GEN, OTB_TRACE_THREADS_FULL>(a, stream, sm_count, smem);
    else {
        if (numeric) return launch_store_threads<POL, CAPS, GEN, OTB_TRACE_THREADS_FULL>(a, stream, sm_count, smem);
        return launch_store_threads<POL, CAPS, GEN, OTB_TRACE_THREADS_FULL_FLAT>(a, stream, sm_count, smem);
    }
}

static int launch_trace_store_range(const OtbScene* scene, TraceArgs& a, cudaStream_t stream, int sm_count)
{
    // aux tables in shared memory (TMA bulk copy at kernel start) for the numeric-surface kernels when they fit
    a.aux_smem_bytes = 0;
    a.pad0 = 0;
    const int64_t aux_bytes = ((scene->n_aux*(int64_t)sizeof(double) + 15)/16)*16;
    if (scene->caps == OTB_CAPS_FULL && scene->n_aux > 0 && aux_bytes <= OTB_AUX_SMEM_MAX) a.aux_smem_bytes = (int)aux_bytes;
    size_t smem = (size_t)a.aux_smem_bytes + 16 + sizeof(int)*OTB_NMSG*a.out.nt;
#define OTB_LAUNCH_STORE(POL, CAPS) do { \
        const int rc_ = (a.G.nsrc > 0) ? launch_store_caps<POL, CAPS, true>(a, stream, sm_count, smem, numeric) \
                                       : launch_store_caps<POL, CAPS, false>(a, stream, sm_count, smem, numeric); \
        if (rc_) return rc_; } while (0)
    bool numeric = false;
    for (int i = 0; i < scene->k.n_steps; ++i) {
        const KSurface& S = scene->k.surf[scene->k.steps[i].surface];
        if (!(S.flags & OTB_SF_FLAT) && (S.kind == OTB_SURF_FUNC || S.kind == OTB_SURF_DATA || S.kind == OTB_SURF_ASPHERE)) numeric = true;
    }
#if OTB_SPEC
    if (!otb_scene_equal(scene->k, K_SPEC_HOST)) {
        otb_set_error("this engine build is specialised for a different scene");
        return OTB_ERR_INVALID_ARG;
    }
    OTB_LAUNCH_STORE((OTB_SPEC_POL != 0), OTB_SPEC_CAPS);
#else
    const bool lean = scene->caps == OTB_CAPS_LENS;
    if (scene->k.no_pol) {
        if (lean) OTB_LAUNCH_STORE(false, OTB_CAPS_LENS);
        else OTB_LAUNCH_STORE(false, OTB_CAPS_FULL);
    } else {
        if (lean) OTB_LAUNCH_STORE(true, OTB_CAPS_LENS);
        else OTB_LAUNCH_STORE(true, OTB_CAPS_FULL);
    }
#endif
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return otb_cuda_fail(e, "trace_store_kernel launch");
    return OTB_OK;
}
