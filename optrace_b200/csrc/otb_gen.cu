// otb_gen.cu — on-device ray generation: RaySource.create_rays (ray_source.py:204-437) with the
// sampling primitives of random.py (stratified grids + shuffle, Shirley disc map, inverse-CDF tables),
// driven by counter-based Philox4x32-10 instead of numpy's SFC64 stream.
//
// Stratification: the reference draws "grid cell + dither" and then shuffles globally (random.py:25-45).
// Here ray m of a source takes grid cell perm_key(m), a keyed Feistel bijection of [0, n) (otb_rng.cuh), so
// every cell is used exactly once and different random variables are decorrelated by different keys.
// Generated bundles therefore agree with the reference statistically (same distributions, same
// stratification), not bit-wise; parity tests inject reference-generated bundles instead.
#include "otb_gen.cuh"

struct GenArgs {
    GenBlock G;
    int64_t N, k_begin, k_end;
    double *p0, *s0;
    float *pol0, *w0, *wl0;
    int* status;
};

#ifndef OTB_GEN_MINBLOCKS
#define OTB_GEN_MINBLOCKS 5      // measured: 6 / 8 resident blocks (80 / 64 registers) are not faster
#endif
__global__ void __launch_bounds__(128, OTB_GEN_MINBLOCKS)
generate_kernel(const __grid_constant__ GenArgs a)
{
    const int64_t N = a.N;
    double* __restrict__ p0 = a.p0;
    double* __restrict__ s0 = a.s0;
    float* __restrict__ pol0 = a.pol0;
    float* __restrict__ w0 = a.w0;
    float* __restrict__ wl0 = a.wl0;
    for (int64_t k = a.k_begin + (int64_t)blockIdx.x*blockDim.x + threadIdx.x; k < a.k_end; k += (int64_t)gridDim.x*blockDim.x) {
        GenRay r;
        generate_ray(a.G, k, r);
        if (r.neg_dir) atomicOr(a.status, OTB_STATUS_NEG_DIR);
        if (!a.G.no_pol) {
            pol0[k] = r.pol[0];
            pol0[k + N] = r.pol[1];
            pol0[k + 2*N] = r.pol[2];
        }
        p0[k] = r.p.x;
        p0[k + N] = r.p.y;
        p0[k + 2*N] = r.p.z;
        s0[k] = r.s.x;
        s0[k + N] = r.s.y;
        s0[k + 2*N] = r.s.z;
        w0[k] = r.w;
        wl0[k] = r.wl;
    }
}

int otb_sm_count();

// validation shared with the fused entry points: the source blocks must tile [0, N) (RayStorage.B_list)
int otb_check_sources(const OtbSource* sources_h, int n_sources, int64_t N)
{
    if (!sources_h || n_sources < 1) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    int64_t cover = 0;
    for (int i = 0; i < n_sources; ++i) {
        if (sources_h[i].ray_start != cover || sources_h[i].n_rays < 0) {
            otb_set_error("sources must cover [0, N) with contiguous blocks (RayStorage.B_list)");
            return OTB_ERR_INVALID_ARG;
        }
        if (sources_h[i].orientation == OTB_OR_FUNCTION && !OTB_HAS_USER_FUNCS) {
            otb_set_error("orientation function without a compiled device function in this engine build");
            return OTB_ERR_UNSUPPORTED;
        }
        cover += sources_h[i].n_rays;
    }
    if (cover != N) { otb_set_error("source ray counts do not sum to N"); return OTB_ERR_INVALID_ARG; }
    return OTB_OK;
}

// source group [g0, g0 + nsrc) of a generator description as a kernel-parameter block
void otb_fill_genblock(GenBlock* G, const OtbSource* sources_h, int g0, int nsrc, const double* gen_aux_d, uint64_t seed,
                       int64_t ray_offset, int no_pol)
{
    memset(G, 0, sizeof(*G));
    G->nsrc = nsrc;
    for (int i = 0; i < nsrc; ++i) {
        G->src[i] = sources_h[g0 + i];
        gen_source_setup(G->src[i], G->sc[i]);
    }
    G->no_pol = no_pol;
    G->src_index0 = g0;
    G->aux = gen_aux_d;
    G->ray_offset = ray_offset;
    G->seed = seed;
}

extern "C" int otb_generate_rays(const OtbSource* sources_h, int n_sources, const double* gen_aux_d, int64_t N,
                                 uint64_t seed, int64_t ray_offset, int no_pol, double* p0_d, double* s0_d,
                                 float* pol0_d, float* w0_d, float* wl_d, int32_t* status_d, void* stream)
{
    if (!sources_h || n_sources < 1 || !p0_d || !s0_d || !w0_d || !wl_d || (!no_pol && !pol0_d) || !status_d) {
        otb_set_error("null argument");
        return OTB_ERR_INVALID_ARG;
    }
    if (N <= 0) return OTB_OK;
    if (int rc = otb_check_sources(sources_h, n_sources, N)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // one launch per group of <= OTB_GEN_MAXSRC sources over the ray range that group covers; no allocation, no sync
    for (int g0 = 0; g0 < n_sources; g0 += OTB_GEN_MAXSRC) {
        GenArgs a;
        const int nsrc = (n_sources - g0 < OTB_GEN_MAXSRC) ? n_sources - g0 : OTB_GEN_MAXSRC;
        otb_fill_genblock(&a.G, sources_h, g0, nsrc, gen_aux_d, seed, ray_offset, no_pol);
        a.N = N;
        a.k_begin = sources_h[g0].ray_start;
        a.k_end = sources_h[g0 + nsrc - 1].ray_start + sources_h[g0 + nsrc - 1].n_rays;
        a.p0 = p0_d; a.s0 = s0_d; a.pol0 = pol0_d; a.w0 = w0_d; a.wl0 = wl_d;
        a.status = status_d;
        const int64_t n = a.k_end - a.k_begin;
        if (n <= 0) continue;
        const int blocks = otb_one_wave_grid(generate_kernel, 128, 0, otb_sm_count(), (n + 127)/128);
        generate_kernel<<<blocks, 128, 0, st>>>(a);
        OTB_CUDA(cudaGetLastError());
    }
    return OTB_OK;
}

// The standard normal deviates the trace kernels draw for HURB aperture `slot` (otb_trace.cu store_step): Philox
// counter = global ray id, stream 0x48555242 ("HURB"), Box-Muller.  Lets a host reproduce a device-RNG trace.
__global__ void __launch_bounds__(256) hurb_normals_kernel(int64_t N, uint64_t seed, int64_t ray_offset, int slot,
                                                           double* __restrict__ za, double* __restrict__ zb)
{
    for (int64_t k = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x*blockDim.x) {
        Philox4 rnd = philox4x32_10((uint64_t)(ray_offset + k), 0x48555242u, (uint32_t)slot, seed);
        double a, b;
        normal2(rnd, a, b);
        za[k] = a;
        zb[k] = b;
    }
}

extern "C" int otb_hurb_normals(int64_t N, uint64_t seed, int64_t ray_offset, int32_t slot, double* za_d, double* zb_d, void* stream)
{
    if (!za_d || !zb_d || N < 0 || slot < 0) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    if (N == 0) return OTB_OK;
    const int64_t b = (N + 255)/256, cap = 8LL*otb_sm_count();
    hurb_normals_kernel<<<(int)(b < cap ? b : cap), 256, 0, (cudaStream_t)stream>>>(N, seed, ray_offset, slot, za_d, zb_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}
