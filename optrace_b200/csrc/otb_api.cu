// otb_api.cu — C ABI of the engine (include/otb.h): lifecycle, error plumbing, scene upload and the
// array-evaluation entry points.  Kernels for trace / detector / generation live in their own files.
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <vector>
#include "otb_common.cuh"
#include "otb_surfaces.cuh"
#include "otb_media.cuh"

// launchers implemented in the other translation units
int otb_launch_trace_store(const OtbScene*, const OtbRays*, const OtbRayStore*, int64_t*, int32_t*, cudaStream_t, int);

static thread_local char g_err[1024] = "";
static int g_sm_count = 148;
static int g_device = -1;

void otb_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int otb_cuda_fail(cudaError_t e, const char* what)
{
    otb_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return e == cudaErrorMemoryAllocation ? OTB_ERR_OOM : OTB_ERR_CUDA;
}

int otb_sm_count() { return g_sm_count; }

extern "C" {

const char* otb_last_error(void) { return g_err; }
int otb_abi_version(void) { return OTB_ABI_VERSION; }

int otb_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        otb_set_error("no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(e));
        return OTB_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        otb_set_error("invalid device %d (have %d)", device, n);
        return OTB_ERR_INVALID_ARG;
    }
    cudaDeviceProp prop;
    OTB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        otb_set_error("device %d (%s) is sm_%d%d; this build targets sm_100a (B200) only", device, prop.name, prop.major, prop.minor);
        return OTB_ERR_UNSUPPORTED;
    }
    OTB_CUDA(cudaSetDevice(device));
    g_sm_count = prop.multiProcessorCount;
    g_device = device;
    return OTB_OK;
}

int otb_device_info(OtbDeviceInfo* out)
{
    if (!out) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    int dev = 0;
    OTB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    OTB_CUDA(cudaGetDeviceProperties(&prop, dev));
    memset(out, 0, sizeof(*out));
    out->device = dev;
    out->sm_major = prop.major;
    out->sm_minor = prop.minor;
    out->sm_count = prop.multiProcessorCount;
    out->total_mem = (int64_t)prop.totalGlobalMem;
    out->l2_bytes = prop.l2CacheSize;
    out->max_smem_per_block = (int32_t)prop.sharedMemPerBlockOptin;
    strncpy(out->name, prop.name, sizeof(out->name) - 1);
    return OTB_OK;
}

int otb_dev_alloc(void** ptr_d, size_t bytes)
{
    if (!ptr_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    OTB_CUDA(cudaMalloc(ptr_d, bytes ? bytes : 1));
    return OTB_OK;
}
int otb_dev_free(void* ptr_d) { OTB_CUDA(cudaFree(ptr_d)); return OTB_OK; }
int otb_memcpy_h2d(void* dst_d, const void* src_h, size_t bytes, void* stream)
{
    OTB_CUDA(cudaMemcpyAsync(dst_d, src_h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return OTB_OK;
}
int otb_memcpy_d2h(void* dst_h, const void* src_d, size_t bytes, void* stream)
{
    OTB_CUDA(cudaMemcpyAsync(dst_h, src_d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return OTB_OK;
}
int otb_memset_d(void* dst_d, int value, size_t bytes, void* stream)
{
    OTB_CUDA(cudaMemsetAsync(dst_d, value, bytes, (cudaStream_t)stream));
    return OTB_OK;
}
int otb_stream_sync(void* stream) { OTB_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return OTB_OK; }

// ---- scene ---------------------------------------------------------------------------------------
static bool scene_needs_user_funcs(const OtbSceneDesc* d)
{
    for (int i = 0; i < d->n_surfaces; ++i) if (d->surfaces[i].kind == OTB_SURF_FUNC) return true;
    for (int i = 0; i < d->n_media; ++i) if (d->media[i].model == OTB_N_FUNCTION) return true;
    for (int i = 0; i < d->n_filters; ++i) if (d->filters[i].type == OTB_T_FUNCTION) return true;
    return false;
}

static int check_scene_desc(const OtbSceneDesc* d)
{
    if (d->abi_version != OTB_ABI_VERSION) { otb_set_error("ABI version mismatch"); return OTB_ERR_INVALID_ARG; }
    if (d->n_steps < 1 || d->n_surfaces < 1 || d->n_media < 1) { otb_set_error("empty scene"); return OTB_ERR_INVALID_ARG; }
    for (int i = 0; i < d->n_steps; ++i) {
        const OtbStep& s = d->steps[i];
        if (s.surface < 0 || s.surface >= d->n_surfaces || s.medium_after >= d->n_media || s.filter >= d->n_filters
            || (s.role == OTB_STEP_FILTER && s.filter < 0) || (s.role <= OTB_STEP_IDEAL_LENS && s.medium_after < 0)) {
            otb_set_error("step %d references an invalid surface/medium/filter", i);
            return OTB_ERR_INVALID_ARG;
        }
    }
    if (scene_needs_user_funcs(d) && !OTB_HAS_USER_FUNCS) {
        otb_set_error("scene contains user callables but this engine build has no compiled user functions");
        return OTB_ERR_UNSUPPORTED;
    }
    if (d->n_steps > OTB_MAX_STEPS || d->n_surfaces > OTB_MAX_STEPS || d->n_media > OTB_MAX_MEDIA || d->n_filters > OTB_MAX_FILTERS) {
        otb_set_error("scene too large for the kernel-parameter scene: at most %d tracing surfaces, %d media, %d filters",
                      OTB_MAX_STEPS, OTB_MAX_MEDIA, OTB_MAX_FILTERS);
        return OTB_ERR_UNSUPPORTED;
    }
    return OTB_OK;
}

// host-side kernel-parameter scene (KScene) and capability level from a descriptor
static void fill_kscene(const OtbSceneDesc* d, OtbScene* sc)
{
    KScene& k = sc->k;
    memset(&k, 0, sizeof(k));
    k.n_steps = d->n_steps;
    k.n_media = d->n_media;
    k.n_filters = d->n_filters;
    k.no_pol = d->no_pol;
    k.medium0 = d->medium0;
    k.n_hurb = d->n_hurb;
    k.arithmetic = d->arithmetic;
    for (int i = 0; i < 6; ++i) k.outline[i] = d->outline[i];
    k.hurb_factor = d->hurb_factor;
    for (int i = 0; i < d->n_steps; ++i) k.steps[i] = d->steps[i];
    for (int i = 0; i < d->n_surfaces; ++i) k.surf[i] = otb_ksurface(d->surfaces[i]);
    for (int i = 0; i < d->n_media; ++i) k.media[i] = d->media[i];
    for (int i = 0; i < d->n_filters; ++i) k.filters[i] = d->filters[i];
    sc->nt = d->n_steps + 1;
    sc->caps = OTB_CAPS_LENS;
    for (int i = 0; i < d->n_surfaces; ++i) {
        const int kd = d->surfaces[i].kind;
        if (kd == OTB_SURF_TILTED || kd == OTB_SURF_ASPHERE || kd == OTB_SURF_FUNC || kd == OTB_SURF_DATA) sc->caps = OTB_CAPS_FULL;
    }
    for (int i = 0; i < d->n_steps; ++i) if (d->steps[i].hurb) sc->caps = OTB_CAPS_FULL;
}

int otb_scene_create(const OtbSceneDesc* d, OtbScene** out)
{
    if (!d || !out) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (int rc = check_scene_desc(d)) return rc;
    OtbScene* sc = new OtbScene();
    memset(sc, 0, sizeof(*sc));
    fill_kscene(d, sc);
    const size_t naux = (d->n_aux > 0 ? (size_t)d->n_aux : 1) + 1;      // + 1: bulk copies move 16-byte multiples
    cudaError_t e = cudaMalloc(&sc->aux_d, sizeof(double)*naux);
    if (e != cudaSuccess) { delete sc; return otb_cuda_fail(e, "cudaMalloc(scene aux)"); }
    if (d->n_aux > 0) {
        e = cudaMemcpy(sc->aux_d, d->aux, sizeof(double)*d->n_aux, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(sc->aux_d); delete sc; return otb_cuda_fail(e, "cudaMemcpy(scene aux)"); }
    }
    sc->n_aux = d->n_aux;
    sc->k.aux = sc->aux_d;
    *out = sc;
    return OTB_OK;
}

int otb_scene_update(OtbScene* sc, const OtbSceneDesc* d, void* stream)
{
    if (!sc || !d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (int rc = check_scene_desc(d)) return rc;
    if (d->n_aux != sc->n_aux) { otb_set_error("aux table size changed: re-create the scene"); return OTB_ERR_INVALID_ARG; }
    fill_kscene(d, sc);
    sc->k.aux = sc->aux_d;
    if (d->n_aux > 0)
        OTB_CUDA(cudaMemcpyAsync(sc->aux_d, d->aux, sizeof(double)*d->n_aux, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return OTB_OK;
}

int otb_scene_destroy(OtbScene* scene)
{
    if (!scene) return OTB_OK;
    cudaFree(scene->aux_d);
    delete scene;
    return OTB_OK;
}

int otb_trace_store(const OtbScene* scene, const OtbRays* rays, const OtbRayStore* out,
                    int64_t* msgs_d, int32_t* status_d, void* stream)
{
    if (!scene || !rays || !out || !msgs_d || !status_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (rays->N != out->N || out->nt != scene->nt) {
        otb_set_error("ray store shape (N=%lld, nt=%d) does not match rays (N=%lld) / scene (nt=%d)",
                      (long long)out->N, out->nt, (long long)rays->N, scene->nt);
        return OTB_ERR_INVALID_ARG;
    }
    if ((!rays->gen_h && (!rays->p0_d || !rays->s0_d || !rays->w0_d || !rays->wl_d || (!scene->k.no_pol && !rays->pol0_d)))
        || !out->p_d || !out->s_d || !out->w_d || !out->n_d || !out->wl_d || (!scene->k.no_pol && !out->pol_d)) {
        otb_set_error("missing ray array");
        return OTB_ERR_INVALID_ARG;
    }
    return otb_launch_trace_store(scene, rays, out, msgs_d, status_d, (cudaStream_t)stream, g_sm_count);
}

// ---- stand-alone array evaluation ----------------------------------------------------------------
struct SurfEvalArgs {
    KSurface S;
    const double* aux;
    int64_t N;
};

__global__ void find_hit_kernel(SurfEvalArgs a, const double* __restrict__ p, const double* __restrict__ s,
                                double* __restrict__ ph, uint8_t* __restrict__ hit, uint8_t* __restrict__ ill, int* status)
{
    int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    V3 P = v3(p[i], p[i + a.N], p[i + 2*a.N]), Sd = v3(s[i], s[i + a.N], s[i + 2*a.N]);
    HitResult h = surf_find_hit<OTB_CAPS_FULL>(a.S, a.aux, P, Sd, status);
    ph[i] = h.p.x;
    ph[i + a.N] = h.p.y;
    ph[i + 2*a.N] = h.p.z;
    hit[i] = h.hit;
    ill[i] = h.ill;
}

__global__ void normals_kernel(SurfEvalArgs a, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ n)
{
    int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    V3 v = surf_normal<OTB_CAPS_FULL>(a.S, a.aux, x[i], y[i]);
    n[i] = v.x;
    n[i + a.N] = v.y;
    n[i + 2*a.N] = v.z;
}

__global__ void values_kernel(SurfEvalArgs a, const double* __restrict__ x, const double* __restrict__ y,
                              double* __restrict__ z, uint8_t* __restrict__ m)
{
    int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    if (z) z[i] = surf_values<OTB_CAPS_FULL>(a.S, a.aux, x[i], y[i]);
    if (m) m[i] = surf_mask(a.S, x[i], y[i]);
}

__global__ void medium_kernel(OtbMedium M, const double* aux, int64_t N, const double* __restrict__ wl, double* __restrict__ n)
{
    int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < N) n[i] = medium_n(M, aux, wl[i]);
}

__global__ void projection_kernel(KSurface S, int method, int64_t N, const double* __restrict__ p, double* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    double x = p[i], y = p[i + N], z = p[i + 2*N];
    sphere_project(S, method, x, y, z);
    out[i] = x;
    out[i + N] = y;
    out[i + 2*N] = z;
}

static int upload_aux(const double* aux_h, int64_t naux, double** aux_d)
{
    *aux_d = nullptr;
    if (naux <= 0) return OTB_OK;
    OTB_CUDA(cudaMalloc(aux_d, sizeof(double)*naux));
    OTB_CUDA(cudaMemcpy(*aux_d, aux_h, sizeof(double)*naux, cudaMemcpyHostToDevice));
    return OTB_OK;
}

static int finish_status(int* status_d, cudaStream_t st)
{
    int status = 0;
    OTB_CUDA(cudaMemcpyAsync(&status, status_d, sizeof(int), cudaMemcpyDeviceToHost, st));
    OTB_CUDA(cudaStreamSynchronize(st));
    cudaFree(status_d);
    if (status & OTB_STATUS_TIMEOUT) {
        otb_set_error("Timeout after 200 iterations in hit finding.");
        return OTB_ERR_NUMERIC_TIMEOUT;
    }
    return OTB_OK;
}

static int check_user(const OtbSurface* S)
{
    if (S->kind == OTB_SURF_FUNC && !OTB_HAS_USER_FUNCS) {
        otb_set_error("FunctionSurface needs an engine build with compiled user functions");
        return OTB_ERR_UNSUPPORTED;
    }
    return OTB_OK;
}

int otb_surface_find_hit(const OtbSurface* surf_h, const double* aux_h, int64_t naux, int64_t N,
                         const double* p_d, const double* s_d, double* ph_d, uint8_t* hit_d, uint8_t* ill_d, void* stream)
{
    if (!surf_h || !p_d || !s_d || !ph_d || !hit_d || !ill_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (int rc = check_user(surf_h)) return rc;
    if (N <= 0) return OTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SurfEvalArgs a;
    a.S = otb_ksurface(*surf_h);
    a.N = N;
    double* aux_d;
    if (int rc = upload_aux(aux_h, naux, &aux_d)) return rc;
    a.aux = aux_d;
    int* status_d;
    OTB_CUDA(cudaMalloc(&status_d, sizeof(int)));
    OTB_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int), st));
    find_hit_kernel<<<(unsigned)((N + 127)/128), 128, 0, st>>>(a, p_d, s_d, ph_d, hit_d, ill_d, status_d);
    OTB_CUDA(cudaGetLastError());
    int rc = finish_status(status_d, st);
    cudaFree(aux_d);
    return rc;
}

int otb_surface_normals(const OtbSurface* surf_h, const double* aux_h, int64_t naux, int64_t N,
                        const double* x_d, const double* y_d, double* n_d, void* stream)
{
    if (!surf_h || !x_d || !y_d || !n_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (int rc = check_user(surf_h)) return rc;
    if (N <= 0) return OTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SurfEvalArgs a;
    a.S = otb_ksurface(*surf_h);
    a.N = N;
    double* aux_d;
    if (int rc = upload_aux(aux_h, naux, &aux_d)) return rc;
    a.aux = aux_d;
    normals_kernel<<<(unsigned)((N + 127)/128), 128, 0, st>>>(a, x_d, y_d, n_d);
    OTB_CUDA(cudaGetLastError());
    OTB_CUDA(cudaStreamSynchronize(st));
    cudaFree(aux_d);
    return OTB_OK;
}

int otb_surface_values(const OtbSurface* surf_h, const double* aux_h, int64_t naux, int64_t N,
                       const double* x_d, const double* y_d, double* z_d, uint8_t* mask_d, void* stream)
{
    if (!surf_h || !x_d || !y_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (int rc = check_user(surf_h)) return rc;
    if (N <= 0) return OTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SurfEvalArgs a;
    a.S = otb_ksurface(*surf_h);
    a.N = N;
    double* aux_d;
    if (int rc = upload_aux(aux_h, naux, &aux_d)) return rc;
    a.aux = aux_d;
    values_kernel<<<(unsigned)((N + 127)/128), 128, 0, st>>>(a, x_d, y_d, z_d, mask_d);
    OTB_CUDA(cudaGetLastError());
    OTB_CUDA(cudaStreamSynchronize(st));
    cudaFree(aux_d);
    return OTB_OK;
}

int otb_medium_eval(const OtbMedium* med_h, const double* aux_h, int64_t naux, int64_t N,
                    const double* wl_d, double* n_d, void* stream)
{
    if (!med_h || !wl_d || !n_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (med_h->model == OTB_N_FUNCTION && !OTB_HAS_USER_FUNCS) {
        otb_set_error("RefractionIndex('Function') needs an engine build with compiled user functions");
        return OTB_ERR_UNSUPPORTED;
    }
    if (N <= 0) return OTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    double* aux_d;
    if (int rc = upload_aux(aux_h, naux, &aux_d)) return rc;
    medium_kernel<<<(unsigned)((N + 127)/128), 128, 0, st>>>(*med_h, aux_d, N, wl_d, n_d);
    OTB_CUDA(cudaGetLastError());
    OTB_CUDA(cudaStreamSynchronize(st));
    cudaFree(aux_d);
    return OTB_OK;
}

int otb_sphere_projection(const OtbSurface* surf_h, int method, int64_t N, const double* p_d, double* out_d, void* stream)
{
    if (!surf_h || !p_d || !out_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (method < OTB_PROJ_EQUIDISTANT || method > OTB_PROJ_STEREOGRAPHIC) { otb_set_error("invalid projection"); return OTB_ERR_INVALID_ARG; }
    if (N <= 0) return OTB_OK;
    projection_kernel<<<(unsigned)((N + 127)/128), 128, 0, (cudaStream_t)stream>>>(otb_ksurface(*surf_h), method, N, p_d, out_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

// self-test hook: q_seq = shared-reciprocal division (otb_common.cuh), q_ieee = the compiler's a/b
__global__ void div_selftest_kernel(int64_t N, const double* __restrict__ a, const double* __restrict__ b,
                                    double* __restrict__ q_seq, double* __restrict__ q_ieee)
{
    int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double y = rcp_seq(b[i]);
    q_seq[i] = div_seq(a[i], b[i], y);
    q_ieee[i] = a[i]/b[i];
}

int otb_selftest_division(int64_t N, const double* a_d, const double* b_d, double* q_seq_d, double* q_ieee_d, void* stream)
{
    if (!a_d || !b_d || !q_seq_d || !q_ieee_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (N <= 0) return OTB_OK;
    div_selftest_kernel<<<(unsigned)((N + 255)/256), 256, 0, (cudaStream_t)stream>>>(N, a_d, b_d, q_seq_d, q_ieee_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

}  // extern "C"
