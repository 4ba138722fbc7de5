/*
 * otb.h — C ABI of the B200-native sequential raytracing engine ("optrace on B200").
 *
 * This is the drop-in boundary for the ONE hot path named in BASELINE.json: ray-bundle
 * propagation through the sequential surface list followed by detector-image binning.
 * The reference (drocheam/optrace 1.8.2) has no FFI; its boundary is the public Python API.
 * Each entry point below therefore cites the reference Python routine it replaces
 * (file:line relative to the reference repository).  All pointers named *_d are DEVICE
 * pointers, all pointers named *_h are HOST pointers; plain C types only.
 *
 * Conventions
 *   - every function returns 0 (OTB_OK) or an OtbStatus error code; the message of the
 *     last error of the calling thread is available through otb_last_error().
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - ray arrays are SoA ("planes"): component c of section i of ray r of a (N, nt, 3)
 *     Fortran-ordered array lives at  base[r + N*i + N*nt*c]  — byte-for-byte the layout of
 *     optrace.tracer.ray_storage.RayStorage (ray_storage.py:80-90).
 */
#ifndef OTB_H
#define OTB_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTB_ABI_VERSION 2

typedef enum OtbStatus {
    OTB_OK = 0,
    OTB_ERR_INVALID_ARG = 1,
    OTB_ERR_GEOMETRY = 2,
    OTB_ERR_OOM = 3,
    OTB_ERR_CUDA = 4,
    OTB_ERR_UNSUPPORTED = 5,       /* e.g. user callable without a compiled device function */
    OTB_ERR_NUMERIC_TIMEOUT = 6,   /* regula falsi did not converge in 200 iterations (surface.py:403) */
    OTB_ERR_INDEX_BELOW_ONE = 7    /* refraction index < 1 for a traced wavelength (refraction_index.py:165) */
} OtbStatus;

/* ---- surfaces (optrace/tracer/geometry/surface/*.py) ------------------------------------ */
typedef enum OtbSurfKind {
    OTB_SURF_CIRCLE = 0,   /* circular_surface.py */
    OTB_SURF_RECT = 1,     /* rectangular_surface.py */
    OTB_SURF_RING = 2,     /* ring_surface.py */
    OTB_SURF_SLIT = 3,     /* slit_surface.py */
    OTB_SURF_CONIC = 4,    /* conic_surface.py, spherical_surface.py (k == 0) */
    OTB_SURF_TILTED = 5,   /* tilted_surface.py */
    OTB_SURF_ASPHERE = 6,  /* aspheric_surface.py */
    OTB_SURF_FUNC = 7,     /* function_surface_1d.py / function_surface_2d.py */
    OTB_SURF_DATA = 8      /* data_surface_1d.py / data_surface_2d.py */
} OtbSurfKind;

#define OTB_SF_ROTATED   1   /* rotation angle != 0 (Surface._rotate_rc, surface.py:427) */
#define OTB_SF_1D        2   /* radial profile (FunctionSurface1D / DataSurface1D / Asphere) */
#define OTB_SF_HAS_DERIV 4   /* user derivative function present */
#define OTB_SF_HAS_MASK  8   /* user mask function present */
#define OTB_SF_FLAT      16  /* z_max == z_min (Surface.is_flat, surface.py:47) */
#define OTB_SF_ROTSYM    32  /* rotational_symmetry class attribute */

#define OTB_NPAR 20

/* parameter slots (indices into OtbSurface.par), by kind */
/* RECT / SLIT */
#define OTB_P_DIMX 0
#define OTB_P_DIMY 1
#define OTB_P_COSM 2   /* cos(-angle) */
#define OTB_P_SINM 3   /* sin(-angle) */
#define OTB_P_COSP 4   /* cos(+angle) */
#define OTB_P_SINP 5   /* sin(+angle) */
#define OTB_P_DIMIX 6
#define OTB_P_DIMIY 7
/* RING */
#define OTB_P_RI 0
/* CONIC / ASPHERE */
#define OTB_P_K 0
#define OTB_P_RHO 1        /* 1/R */
#define OTB_P_KP1 2        /* k + 1 */
#define OTB_P_INVRHO 3     /* 1/rho */
#define OTB_P_TWOINVRHO 4  /* 2/rho */
#define OTB_P_RHO2 5       /* rho**2 */
#define OTB_P_KP1RHO2 6    /* (k+1)*rho**2 */
#define OTB_P_KRHO2 7      /* k*rho**2 */
#define OTB_P_EDGEZ 8      /* _values(r - N_EPS, 0): radially continued edge value (surface.py:162) */
#define OTB_P_FDEPS 9      /* finite-difference step of Surface.normals (surface.py:266-270) */
#define OTB_P_ZMIN_E 11    /* CONIC: z_min - N_EPS (conic_surface.py:160) */
#define OTB_P_ZMAX_E 12    /* CONIC: z_max + N_EPS */
#define OTB_P_RB2 13       /* CONIC: (r + N_EPS)^2 (Surface.mask, surface.py:235-245) */
/* TILTED */
#define OTB_P_NX 0
#define OTB_P_NY 1
#define OTB_P_NZ 2
#define OTB_P_MX 3         /* -nx/nz */
#define OTB_P_MY 4         /* -ny/nz */
/* FUNC / DATA: sign, offset + rotation slots 2..5 shared with RECT, EDGEZ, FDEPS shared */
#define OTB_P_SIGN 0
#define OTB_P_OFFSET 1

typedef struct OtbSurface {
    int32_t kind;       /* OtbSurfKind */
    int32_t flags;      /* OTB_SF_* */
    int32_t func_id;    /* user device-function slot for OTB_SURF_FUNC, else -1 */
    int32_t aux_off;    /* offset (in doubles) of this surface's table in the scene aux buffer */
    int32_t aux_n0;     /* ASPHERE: number of polyval coefficients; DATA: number of x knots */
    int32_t aux_n1;     /* ASPHERE: number of derivative coefficients; DATA: number of y knots (0 = 1-D) */
    int32_t pad0, pad1;
    double pos[3];
    double r;
    double z_min, z_max;
    double par[OTB_NPAR];
} OtbSurface;

/* ---- media (optrace/tracer/refraction_index.py:62-169) ---------------------------------- */
typedef enum OtbMediumModel {
    OTB_N_CONSTANT = 0, OTB_N_ABBE = 1, OTB_N_CAUCHY = 2, OTB_N_CONRADY = 3,
    OTB_N_SELLMEIER1 = 4, OTB_N_SELLMEIER2 = 5, OTB_N_SELLMEIER3 = 6, OTB_N_SELLMEIER4 = 7,
    OTB_N_SELLMEIER5 = 8, OTB_N_SCHOTT = 9, OTB_N_HERZBERGER = 10, OTB_N_HANDBOOK1 = 11,
    OTB_N_HANDBOOK2 = 12, OTB_N_EXTENDED = 13, OTB_N_EXTENDED2 = 14, OTB_N_EXTENDED3 = 15,
    OTB_N_DATA = 16,      /* np.interp table (wls, vals) in aux */
    OTB_N_FUNCTION = 17   /* user device function */
} OtbMediumModel;

typedef struct OtbMedium {
    int32_t model;
    int32_t func_id;
    int32_t aux_off;   /* DATA: wls at aux_off, vals at aux_off + aux_n */
    int32_t aux_n;
    double c[12];      /* CONSTANT: c[0]=n; ABBE: c[0]=A, c[1]=B, c[2]=d; others: coeff list */
} OtbMedium;

/* ---- filter spectra (spectrum.py:81-119, transmission_spectrum.py:73-84) ----------------- */
typedef enum OtbSpectrumType {
    OTB_T_CONSTANT = 0, OTB_T_DATA = 1, OTB_T_RECTANGLE = 2, OTB_T_GAUSSIAN = 3, OTB_T_FUNCTION = 4
} OtbSpectrumType;

typedef struct OtbFilter {
    int32_t type;
    int32_t inverse;   /* TransmissionSpectrum.inverse */
    int32_t func_id;
    int32_t aux_off, aux_n, pad;
    double c[4];       /* CONSTANT: val; RECTANGLE: wl0, wl1, val; GAUSSIAN: val, mu, sig */
} OtbFilter;

/* ---- sequential steps = tracing surfaces (raytracer.py:307-397) -------------------------- */
typedef enum OtbStepRole {
    OTB_STEP_LENS_FRONT = 0, OTB_STEP_LENS_BACK = 1, OTB_STEP_IDEAL_LENS = 2,
    OTB_STEP_FILTER = 3, OTB_STEP_APERTURE = 4
} OtbStepRole;

typedef struct OtbStep {
    int32_t role;
    int32_t surface;      /* index into the surface array */
    int32_t medium_after; /* medium of the section after this surface, -1 = unchanged */
    int32_t filter;       /* filter index for OTB_STEP_FILTER */
    int32_t hurb;         /* 1: aperture bends rays (use_hurb and not the end absorber), raytracer.py:385 */
    int32_t hurb_slot;    /* index of this aperture among the HURB apertures (normal-deviate stream) */
    double D;             /* optical power of an ideal lens (raytracer.py:748) */
} OtbStep;

typedef struct OtbSceneDesc {
    int32_t abi_version;
    int32_t n_surfaces, n_steps, n_media, n_filters;
    int32_t no_pol;        /* Raytracer.no_pol */
    int32_t medium0;       /* ambient medium n0 */
    int32_t n_hurb;        /* number of bending apertures */
    int64_t n_aux;
    double outline[6];     /* Raytracer.outline */
    double hurb_factor;    /* Raytracer.HURB_FACTOR */
    const OtbSurface* surfaces;
    const OtbStep* steps;
    const OtbMedium* media;
    const OtbFilter* filters;
    const double* aux;
    int32_t arithmetic;    /* OTB_ARITH_*: floating-point contract of the lens-surface step */
    int32_t pad;
} OtbSceneDesc;

/* OTB_ARITH_EXACT: every + - * / sqrt rounds like the reference's numpy float64 operation (results bit-identical
 * on closed-form scenes).  OTB_ARITH_RELAXED: fused multiply-adds, reciprocal-multiply division, rsqrt-based
 * square roots and normalisation; results within ~1e-14 of the reference (acceptance criterion: 1e-9). */
#define OTB_ARITH_EXACT 0
#define OTB_ARITH_RELAXED 1

typedef struct OtbScene OtbScene;   /* opaque, device-resident copy of the descriptor */

/* info message rows (Raytracer.INFOS, raytracer.py:43-48) */
#define OTB_MSG_ABSORB_MISSING 0
#define OTB_MSG_TIR 1
#define OTB_MSG_ILL_COND 2
#define OTB_MSG_OUTLINE 3
#define OTB_MSG_HURB_NEG 4
#define OTB_NMSG 5

/* ---- ray bundles -------------------------------------------------------------------------- */
/* Initial rays = output of RaySource.create_rays (ray_source.py:204-437), SoA planes. */
struct OtbGenerator;
typedef struct OtbRays {
    int64_t N;
    const double* p0_d;    /* (N,3) F-order */
    const double* s0_d;    /* (N,3) F-order */
    const float* pol0_d;   /* (N,3) F-order, NULL when no_pol */
    const float* w0_d;     /* (N) */
    const float* wl_d;     /* (N) */
    const double* hurb_z_d; /* optional injected standard normals, shape (n_hurb, 2, N); NULL = Philox */
    uint64_t seed;         /* Philox key for HURB when hurb_z_d == NULL */
    int64_t ray_offset;    /* global id of local ray 0 (multi-GPU shards) */
    const struct OtbGenerator* gen_h;  /* non-NULL: the N rays are GENERATED inside the trace kernel (fused
                              RaySource.create_rays, see OtbGenerator below); p0_d .. wl_d are ignored and may be NULL.
                              A bundle generated this way never exists in HBM: section 0 of the ray store receives
                              its positions / weights / polarisation / wavelengths like any other section. */
} OtbRays;

/* Per-surface ray storage = RayStorage arrays (ray_storage.py:80-90). */
typedef struct OtbRayStore {
    int64_t N;
    int32_t nt;
    int32_t pad;
    double* p_d;     /* (N, nt, 3) F-order */
    double* s_d;     /* (N, 3) F-order: final directions (s0_list aliasing, SURVEY hard part 10) */
    float* pol_d;    /* (N, nt, 3) F-order, NULL when no_pol */
    float* w_d;      /* (N, nt) F-order */
    double* n_d;     /* (N, nt) F-order */
    float* wl_d;     /* (N) */
    const int32_t* trace_status_d;   /* optional: status word of the trace that filled the store (device); when it
                                        lacks OTB_STATUS_Z_DECREASE the detector search may bisect the sections */
} OtbRayStore;

/* ---- detectors (raytracer.py:881-1051, render_image.py:361-421) ------------------------- */
typedef enum OtbProjection {
    OTB_PROJ_NONE = 0, OTB_PROJ_EQUIDISTANT = 1, OTB_PROJ_ORTHOGRAPHIC = 2,
    OTB_PROJ_EQUAL_AREA = 3, OTB_PROJ_STEREOGRAPHIC = 4
} OtbProjection;

typedef struct OtbDetector {
    OtbSurface surface;
    int32_t projection;   /* OtbProjection, only for spherical detector surfaces */
    int32_t has_extent;   /* 1: user extent given (hits outside are dropped, raytracer.py:1034-1040) */
    double extent[4];     /* user extent [x0, x1, y0, y1] */
} OtbDetector;

/* ---- sources for on-device generation (ray_source.py:204-437, random.py) ---------------- */
typedef enum OtbSourceShape {
    OTB_SHAPE_POINT = 0, OTB_SHAPE_LINE = 1, OTB_SHAPE_CIRCLE = 2, OTB_SHAPE_RING = 3,
    OTB_SHAPE_RECT = 4, OTB_SHAPE_IMAGE_RGB = 5, OTB_SHAPE_IMAGE_GRAY = 6
} OtbSourceShape;
typedef enum OtbOrientation { OTB_OR_CONSTANT = 0, OTB_OR_CONVERGING = 1,
    OTB_OR_FUNCTION = 2   /* or_func(x, y) (ray_source.py:274-276) as a compiled device function */
} OtbOrientation;
typedef enum OtbDivergence {
    OTB_DIV_NONE = 0, OTB_DIV_LAMBERTIAN = 1, OTB_DIV_ISOTROPIC = 2, OTB_DIV_FUNCTION = 3
} OtbDivergence;
typedef enum OtbPolarization {
    OTB_POL_CONSTANT = 0,   /* also "x" (0) and "y" (pi/2) */
    OTB_POL_UNIFORM = 1,
    OTB_POL_LIST = 2,       /* also "xy" */
    OTB_POL_FUNCTION = 3
} OtbPolarization;
typedef enum OtbWavelengthMode {
    OTB_WL_MONO = 0,       /* Monochromatic */
    OTB_WL_UNIFORM = 1,    /* Constant / Rectangle: stratified uniform in [wl0, wl1] */
    OTB_WL_DISCRETE = 2,   /* Lines: discrete inverse CDF (random.py:129-140) */
    OTB_WL_CDF = 3,        /* Data / Blackbody / Function / Histogram: continuous inverse CDF (random.py:143-157) */
    OTB_WL_GAUSSIAN = 4,   /* truncated Gaussian via erfinv (light_spectrum.py:109-122) */
    OTB_WL_SRGB = 5        /* RGB image: primary choice + per-primary inverse CDF (srgb.py:513-553) */
} OtbWavelengthMode;

typedef struct OtbSource {
    int32_t shape, orientation, divergence, polarization, wl_mode;
    int32_t div_2d;
    int32_t img_w, img_h;         /* image sources: pixel grid */
    int64_t n_rays;               /* rays of this source in this launch (RayStorage.N_list) */
    int64_t ray_start;            /* first local ray index of this source (RayStorage.B_list) */
    int64_t gid_start;            /* global id of that ray = Philox counter of its draws: a rank of a sharded trace
                                     holds a slice of every source, the slices of all ranks tile the source's block */
    double power;                 /* weight = power / n_rays_total_of_source (float32) */
    double weight;                /* float32 weight of every ray, computed by the host like ray_source.py:219-220 */
    double pos[3];
    double geom[8];               /* CIRCLE: r; RING: ri, r; RECT: dimx, dimy, cos(a), sin(a); LINE: r, cos(a), sin(a) */
    double extent[4];             /* image sources: surface extent x0, x1, y0, y1 */
    double s[3];                  /* constant orientation */
    double conv_pos[3];           /* converging orientation */
    double div_sin;               /* sin(radians(div_angle)) */
    double div_angle;             /* radians(div_angle) */
    double div_axis;              /* radians(div_axis_angle), 2-D divergence */
    double pol_angle;             /* radians, constant polarisation */
    double wl[4];                 /* MONO: wl; UNIFORM: wl0, wl1; GAUSSIAN: mu, sig, Xl, Xr */
    /* tables in the generator aux buffer (offsets in doubles) */
    int32_t wl_tab_off, wl_tab_n;       /* DISCRETE: x then F ; CDF: x then F */
    int32_t div_tab_off, div_tab_n;     /* FUNCTION divergence: x then F */
    int32_t pol_tab_off, pol_tab_n;     /* LIST: angles then F ; FUNCTION: x then F */
    int32_t pix_cdf_off, pix_cdf_n;     /* image: pixel indices (pix_cdf_n doubles) then cumulative pixel power F */
    int32_t pix_rgb_off;                /* RGB image: primary thresholds (r, r+g) per pixel, 2 doubles each */
    int32_t srgb_off;                   /* RGB image: wl[5000], F_r[5000], F_g[5000], F_b[5000] (srgb.py:528-551) */
    int32_t or_func_id;                 /* OTB_OR_FUNCTION: user device-function slot */
    int32_t coherent;                   /* 1: warp-coherent bundle order.  The reference shuffles every stratified sample
                                           (random.py:41-45); with this flag the strata of ONE random variable (the
                                           direction inside the divergence cone, else the position on the source) are
                                           handed out in blocks of 32 neighbouring cells, so the 32 rays of a warp meet
                                           stops and lens edges together.  Same cells, same distribution of the bundle;
                                           only the ORDER of the rays inside a source block is less random. */
} OtbSource;

/* On-device generation fused into otb_trace_store / otb_trace_render (OtbRays.gen_h): the same source records and
 * tables otb_generate_rays takes.  sources_h[i].ray_start / n_rays tile [0, OtbRays.N). */
typedef struct OtbGenerator {
    const OtbSource* sources_h;
    int32_t n_sources;
    int32_t pad;
    const double* gen_aux_d;
} OtbGenerator;

typedef struct OtbDeviceInfo {
    int32_t device, sm_major, sm_minor, sm_count;
    int64_t total_mem;
    int32_t l2_bytes, max_smem_per_block;
    char name[64];
} OtbDeviceInfo;

/* ---- entry points --------------------------------------------------------------------------- */

/* Asynchronous error reporting: kernels OR bits into a caller-provided device word `status_d` (int32, zeroed by
 * the caller); the host reads it when it synchronises anyway and maps it with otb_status_message().  No entry
 * point of the ray path allocates device memory or synchronises the stream. */
#define OTB_STATUS_TIMEOUT 1       /* Illinois hit finder hit its 200-iteration limit (surface.py:403) */
#define OTB_STATUS_NBELOW1 2       /* refraction index < 1 (refraction_index.py:165) */
#define OTB_STATUS_UNSUPPORTED 4
#define OTB_STATUS_Z_DECREASE 16   /* informational: some stored z decreased from one section to the next */
#define OTB_STATUS_NEG_DIR 8       /* generated direction with s_z <= 0 (ray_source.py:353) */

/* Selects the CUDA device for the calling thread and verifies it is sm_100 (no CPU fallback). */
int otb_init(int device);
const char* otb_last_error(void);
int otb_abi_version(void);
int otb_device_info(OtbDeviceInfo* out);

/* Raw device memory helpers for hosts that do not bring their own allocator (torch does). */
int otb_dev_alloc(void** ptr_d, size_t bytes);
int otb_dev_free(void* ptr_d);
int otb_memcpy_h2d(void* dst_d, const void* src_h, size_t bytes, void* stream);
int otb_memcpy_d2h(void* dst_h, const void* src_d, size_t bytes, void* stream);
int otb_memset_d(void* dst_d, int value, size_t bytes, void* stream);
int otb_stream_sync(void* stream);

/* Copies the flattened scene to the device.  Replaces the per-trace Python object walk of
 * Raytracer.__tracing_elements / sub_trace (raytracer.py:297-397, 492-508). */
int otb_scene_create(const OtbSceneDesc* desc, OtbScene** out);
int otb_scene_destroy(OtbScene* scene);
/* Re-sends a descriptor of the same shape (same aux table size) into an existing scene: host record refreshed,
 * aux tables copied asynchronously on `stream`, no allocation and no device synchronisation.  For callers that
 * re-upload the scene at every trace (the reference re-walks its object tree per trace, raytracer.py:297-305). */
int otb_scene_update(OtbScene* scene, const OtbSceneDesc* desc, void* stream);

/* ---- focus search: Raytracer.focus_search (raytracer.py:1354-1640) on device-resident rays ---------------------
 * prepare: per ray of [ray_begin, ray_end) the section before the first stored point behind z and the line
 *   hit(z') = (pax + sbx z', pay + sby z') through it, its float32 weight and a used flag (0: no such section);
 *   n_use_d += number of used rays (raytracer.py:1553-1578, RayStorage.rays_by_mask ray_storage.py:235-293).
 * moments: weighted sums over the used rays into out_d[4] (zeroed first); par_d = device array [z, ...]:
 *   mode 0: [sum w, sum w^2, sum w x, sum w y] at z              (np.cov / np.average, raytracer.py:1376-1379, 1620)
 *   mode 1: [sum w (x - par[1])^2, sum w (y - par[2])^2] at z
 *   mode 2: the two sums of the direct RMS solution (raytracer.py:1430-1442) with pb0 = par[1:3], v/vz = par[3:5]
 * image: rng_d[4] = range of the hit positions at z, img_d (npx, npx) = weighted histogram over that range
 *   (raytracer.py:1387-1392, misc.binning_indices_2d misc.py:59-91). */
int otb_focus_prepare(const OtbRayStore* store, int64_t ray_begin, int64_t ray_end, double z,
                      double* pax_d, double* pay_d, double* sbx_d, double* sby_d, float* w_d, uint8_t* use_d,
                      int64_t* n_use_d, void* stream);
int otb_focus_moments(const double* pax_d, const double* pay_d, const double* sbx_d, const double* sby_d, const float* w_d,
                      const uint8_t* use_d, int64_t n, int32_t mode, const double* par_d, double* out_d, void* stream);
int otb_focus_image(const double* pax_d, const double* pay_d, const double* sbx_d, const double* sby_d, const float* w_d,
                    const uint8_t* use_d, int64_t n, double z, int32_t npx, int32_t phase, double* rng_d, double* img_d,
                    void* stream);   /* phase 0: range + histogram, 1: range only, 2: histogram over the range in rng_d */

/* ---- spectrum histograms: LightSpectrum.render (light_spectrum.py:40-79) behind Raytracer.detector_spectrum /
 * source_spectrum (raytracer.py:1100-1132, 1311-1329) on device-resident wavelengths and weights -------------
 * stats: count_d[0] += rays used, count_d[1] += rays with non-zero weight among them (the reference sizes the bin
 *   number from np.count_nonzero(w)); range_d[2] (float32, caller initialises to +inf, -inf) = min / max wavelength.
 *   positive_only != 0: only rays with w > 0 are used (the hit selection of _hit_detector, raytracer.py:1023-1024).
 * hist: hist_d[nbins] (float64, accumulated) += weights binned with np.histogram's uniform-bin rule on the float32
 *   edges edges_d[nbins + 1] the host built with np.linspace exactly like numpy does. */
int otb_spectrum_stats(const float* wl_d, const float* w_d, int64_t M, int32_t positive_only, int64_t* count_d,
                       float* range_d, void* stream);
int otb_spectrum_hist(const float* wl_d, const float* w_d, int64_t M, int32_t positive_only, const float* edges_d,
                      int32_t nbins, double* hist_d, void* stream);

/* Resolution-limit filter, RenderImage._apply_rayleigh_filter (render_image.py:255-296): out = max(img (*) psf, 0),
 * zero padded "same" convolution of every XYZW channel with the host-built (K, K) Airy-disc table, K odd.
 * Replaces scipy.signal.fftconvolve of the reference's host path (direct convolution: the kernel is compact). */
int otb_image_convolve(const double* img_d, int32_t Ny, int32_t Nx, const double* psf_d, int32_t K, double* out_d, void* stream);

/* ---- image post-processing: RenderImage.get (render_image.py:131-222) -------------------------------------
 * Join-bins rescaling and per-pixel conversion of the (Ny, Nx, 4) XYZW histogram on the device; replaces
 * cv2.resize(INTER_AREA) + color.xyz_to_srgb / xyz_to_luv / luv_hue / luv_chroma / luv_saturation /
 * outside_srgb_gamut (color/srgb.py:118-407, color/luv.py:20-139) of the reference's host path. */
typedef enum OtbImageMode {
    OTB_IMG_IRRADIANCE = 0, OTB_IMG_ILLUMINANCE = 1, OTB_IMG_SRGB_ABSOLUTE = 2, OTB_IMG_SRGB_PERCEPTUAL = 3,
    OTB_IMG_OUTSIDE_GAMUT = 4, OTB_IMG_LIGHTNESS = 5, OTB_IMG_HUE = 6, OTB_IMG_CHROMA = 7, OTB_IMG_SATURATION = 8
} OtbImageMode;
#define OTB_IMG_NSTATS 8
/* out (Ny/fact, Nx/fact, 4) = block means of img (Ny, Nx, 4); fact must divide both sides (render_image.py:164-174) */
int otb_image_rescale(const double* img_d, int32_t Ny, int32_t Nx, int32_t fact, double* out_d, void* stream);
/* image-wide extrema the conversions normalise by, into stats_d[OTB_IMG_NSTATS] (layout: otb_image.cu).
 * pass 1: everything that depends on the pixels only; pass 2 (param = L_th) and pass 3 (param = chroma_scale):
 * the two statistics of the perceptual rendering intent that depend on earlier ones (srgb.py:238-262, 331-354) */
int otb_image_stats(const double* img_d, int64_t npx, int32_t pass, double param, double* stats_d, void* stream);
/* out (npx) or (npx, 3) for the two sRGB modes.  scale: 1/Apx (irradiance) or K/Apx (illuminance);
 * chroma_scale: perceptual intent only, negative = no colour outside the gamut (plain conversion, srgb.py:308-310) */
int otb_image_convert(const double* img_d, int64_t npx, int32_t mode, double scale, double chroma_scale,
                      const double* stats_d, double* out_d, void* stream);

/* ---- sparse transport of detector images (otb_tiles.cu) --------------------------------------------------------
 * A (Ny, Nx, 4) histogram of an imaging system is a few per cent non-zero; the all-reduce over the GPUs of a
 * sharded trace (the reference has no counterpart) and the device -> host copy behind
 * RenderImage.data move only the occupied T x T pixel tiles:
 *   mask:   mask_d[ty*ntx + tx] = 1 where the tile holds a non-zero value (never cleared: the caller zeroes it, and
 *           MAX-all-reduces it over the ranks to obtain the union)
 *   pack:   header_d[0] = number of tiles in the mask, header_d[1] = 1 when it exceeds cap (nothing else is valid
 *           then), header_d[2 + k] = id of the k-th tile; packed_d[k][T][T][4] = copy of that tile (zero padded at
 *           the image edges, unused slots zeroed)            header_d: int32[2 + cap], packed_d: double[cap*T*T*4]
 *   unpack: the packed tiles written back into the image (no-op on overflow) */
int otb_image_tiles_mask(const double* img_d, int32_t Ny, int32_t Nx, int32_t T, int32_t* mask_d, void* stream);
int otb_image_tiles_pack(const double* img_d, int32_t Ny, int32_t Nx, int32_t T, const int32_t* mask_d, int32_t cap,
                         int32_t* header_d, double* packed_d, void* stream);
int otb_image_tiles_unpack(double* img_d, int32_t Ny, int32_t Nx, int32_t T, const int32_t* header_d, int32_t cap,
                           const double* packed_d, void* stream);

/* Store-mode trace: replaces Raytracer.trace's sub_trace surface loop (raytracer.py:297-397)
 * including find_hit, __refraction, __compute_polarization, __refraction_ideal_lens, __hurb,
 * __outline_intersection, Filter/Aperture handling.  msgs_d: int64[OTB_NMSG * nt], accumulated. */
int otb_trace_store(const OtbScene* scene, const OtbRays* rays, const OtbRayStore* out,
                    int64_t* msgs_d, int32_t* status_d, void* stream);

/* On-device ray generation: replaces RaySource.create_rays (ray_source.py:204-437) and the
 * sampling primitives of random.py with counter-based Philox4x32-10. */
int otb_generate_rays(const OtbSource* sources_h, int n_sources, const double* gen_aux_d,
                      int64_t N, uint64_t seed, int64_t ray_offset, int no_pol,
                      double* p0_d, double* s0_d, float* pol0_d, float* w0_d, float* wl_d,
                      int32_t* status_d, void* stream);

/* The two standard normal deviates per ray that otb_trace_store / otb_trace_render draw for HURB aperture `slot`
 * when OtbRays.hurb_z_d is NULL (stand-ins for np.random.normal in Raytracer.__hurb, raytracer.py:468-469): Philox
 * counter = ray_offset + ray, key = seed.  za_d, zb_d: double[N].  For hosts that replay a device-RNG trace. */
int otb_hurb_normals(int64_t N, uint64_t seed, int64_t ray_offset, int32_t slot, double* za_d, double* zb_d, void* stream);

/* Detector hits from stored sections: replaces Raytracer._hit_detector (raytracer.py:881-1051)
 * up to and including the sphere projection.  Outputs per ray: projected hit x, y (double),
 * weight (float, 0 = no valid hit).  range_d: double[4] = min x, max x, min y, max y over valid
 * hits (atomically merged, caller initialises to +inf,-inf,+inf,-inf). ill_d: int64 counter. */
int otb_detector_hits(const OtbRayStore* store, int64_t ray_begin, int64_t ray_end,
                      const OtbDetector* det_h, double* hx_d, double* hy_d, float* hw_d,
                      double* range_d, int64_t* ill_d, int32_t* status_d, void* stream);

/* Histogram binning: replaces RenderImage.render's binning_indices_2d + observer weighting +
 * np.add.at (render_image.py:390-417, misc.py:59-91, observers.py:14-41).
 * img_d: double (Ny, Nx, 4) C-order [X, Y, Z, W], accumulated; cnt_d: optional int32 (Ny, Nx)
 * hit counts (NULL to skip). */
int otb_render_xyzw(const double* x_d, const double* y_d, const float* w_d, const float* wl_d,
                    int64_t M, const double extent[4], int32_t Nx, int32_t Ny,
                    double* img_d, int32_t* cnt_d, void* stream);

/* Fused render mode: generation-free trace + detector test + binning with no per-surface
 * storage — the per-chunk body of Raytracer.iterative_render (raytracer.py:1235-1264).
 * Bin mode (img_d != NULL): for every detector k the frozen extent extents_h[4*k..] (raytracer.py:1262) and
 * the grid Nx_h[k] x Ny_h[k] must be given; img_d[k] is a (Ny, Nx, 4) double image, cnt_d[k] an optional
 * (Ny, Nx) int32 count image.  Range mode (img_d == NULL): nothing is binned, range_d[4*k..] receives
 * min x, max x, min y, max y of the valid hits (auto extent of the first chunk, raytracer.py:1042-1046). */
int otb_trace_render(const OtbScene* scene, const OtbRays* rays, int n_det,
                     const OtbDetector* dets_h, const double* extents_h,
                     const int32_t* Nx_h, const int32_t* Ny_h,
                     double* const* img_d, int32_t* const* cnt_d, double* range_d,
                     int64_t* msgs_d, int32_t* status_d, void* stream);

/* Stand-alone surface evaluation on arrays (device pointers): Surface.find_hit / normals /
 * values / mask (surface.py:137-164, 235-285, 307-414 and the per-class overrides). */
int otb_surface_find_hit(const OtbSurface* surf_h, const double* aux_h, int64_t naux, int64_t N,
                         const double* p_d, const double* s_d,
                         double* ph_d, uint8_t* hit_d, uint8_t* ill_d, void* stream);
int otb_surface_normals(const OtbSurface* surf_h, const double* aux_h, int64_t naux, int64_t N,
                        const double* x_d, const double* y_d, double* n_d, void* stream);
int otb_surface_values(const OtbSurface* surf_h, const double* aux_h, int64_t naux, int64_t N,
                       const double* x_d, const double* y_d, double* z_d, uint8_t* mask_d, void* stream);

/* Refraction index / filter evaluation on arrays: RefractionIndex.__call__ (refraction_index.py:62). */
int otb_medium_eval(const OtbMedium* med_h, const double* aux_h, int64_t naux, int64_t N,
                    const double* wl_d, double* n_d, void* stream);

/* SphericalSurface.sphere_projection (spherical_surface.py:36-97) on (N,3) F-order points. */
int otb_sphere_projection(const OtbSurface* surf_h, int method, int64_t N, const double* p_d, double* out_d, void* stream);

/* Self-test: the engine's shared-reciprocal division (q_seq) next to the compiler's IEEE division (q_ieee);
 * the two must be bit-identical (parity contract of the vector normalisations). */
int otb_selftest_division(int64_t N, const double* a_d, const double* b_d, double* q_seq_d, double* q_ieee_d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OTB_H */
