// otb_gen.cu — on-device ray generation: RaySource.create_rays (ray_source.py:204-437) with the
// sampling primitives of random.py (stratified grids + shuffle, Shirley disc map, inverse-CDF tables),
// driven by counter-based Philox4x32-10 instead of numpy's SFC64 stream.
//
// Stratification: the reference draws "grid cell + dither" and then shuffles globally (random.py:25-45).
// Here ray m of a source takes grid cell perm_key(m), a keyed Feistel bijection of [0, n) (otb_rng.cuh), so
// every cell is used exactly once and different random variables are decorrelated by different keys.
// Generated bundles therefore agree with the reference statistically (same distributions, same
// stratification), not bit-wise; parity tests inject reference-generated bundles instead.
#include "otb_common.cuh"
#include "otb_rng.cuh"


struct GenCtx {
    uint64_t seed;
    uint64_t gid;      // global ray id (Philox counter)
    uint64_t m;        // index of the ray inside its source
    uint64_t n;        // rays of this source in this launch
    uint32_t src;      // source index (decorrelates sources)
    // per-source constants, computed when the thread moves to another source
    uint32_t fa, fb;   // moduli of the stratum permutation (feistel_setup)
    uint64_t N2;       // floor(sqrt(n)): side of the stratified square grid (random.py:25-33)
    double inv_n;      // 1 / n
    Philox4 A, B;      // shared random blocks of this ray (see strat1)
};

enum { ST_POS = 1, ST_WL = 2, ST_RGB = 3, ST_DIV = 4, ST_DIV2 = 5, ST_POL = 6, ST_PIX = 7, ST_PIXOFF = 8 };

__device__ __forceinline__ Philox4 draw(const GenCtx& g, uint32_t stream) { return philox4x32_10(g.gid, stream, g.src, g.seed); }
__device__ __forceinline__ uint64_t stratum(const GenCtx& g, uint32_t stream)
{
    return feistel_perm_ab(g.m, g.n, g.fa, g.fb, g.seed ^ ((uint64_t)stream << 40) ^ ((uint64_t)g.src << 20) ^ 0x5bd1e995u);
}

// random.stratified_interval_sampling (random.py:48-66): value in [a, b)
// One-dimensional draws need 64 of the 128 bits of a Philox block: wavelength + polarisation angle share block A,
// RGB primary + image pixel share block B (both computed once per ray); the strata stay decorrelated through the
// per-stream keys of the permutation.
__device__ __forceinline__ double strat1(const GenCtx& g, uint32_t stream, double a, double b)
{
    uint32_t hi, lo;
    if (stream == ST_WL) { hi = g.A.v[0]; lo = g.A.v[1]; }
    else if (stream == ST_POL) { hi = g.A.v[2]; lo = g.A.v[3]; }
    else if (stream == ST_RGB) { hi = g.B.v[0]; lo = g.B.v[1]; }
    else if (stream == ST_PIX) { hi = g.B.v[2]; lo = g.B.v[3]; }
    else { Philox4 r = draw(g, stream); hi = r.v[0]; lo = r.v[1]; }
    double dba = (b - a)*g.inv_n;
    return a + ((double)stratum(g, stream) + u01(hi, lo))*dba;
}

// random.stratified_rectangle_sampling (random.py:8-45)
__device__ __forceinline__ void strat2(const GenCtx& g, uint32_t stream, double a, double b, double c, double d, double& x, double& y)
{
    Philox4 r = draw(g, stream);
    double u1 = u01(r.v[0], r.v[1]), u2 = u01(r.v[2], r.v[3]);
    const uint64_t N2 = g.N2;
    uint64_t j = stratum(g, stream);
    if (j < N2*N2) {
        uint64_t iy, ix;
        if (g.n < 0x80000000ull) {          // 32-bit division (the common case) is several times cheaper
            const uint32_t q = (uint32_t)j/(uint32_t)N2;
            iy = q;
            ix = (uint32_t)j - q*(uint32_t)N2;
        } else {
            iy = j/N2;
            ix = j - iy*N2;
        }
        x = a + ((double)ix + u1)*((b - a)/(double)N2);
        y = c + ((double)iy + u2)*((d - c)/(double)N2);
    } else {            // remaining N - N2^2 samples are plain uniform (random.py:36-37)
        x = a + u1*(b - a);
        y = c + u2*(d - c);
    }
}

// random.stratified_ring_sampling (random.py:70-110): Shirley equal-area map + disc->annulus map
__device__ __forceinline__ void strat_ring(const GenCtx& g, uint32_t stream, double ri, double r, bool polar, double& o1, double& o2)
{
    double x, y;
    strat2(g, stream, -r, r, -r, r, x, y);
    double x2 = x*x, y2 = y*y, r_ = 0.0, theta = 0.0;
    if (x2 > y2) {             // theta in units of pi
        r_ = x;
        theta = 0.25*y/x;
    } else if (y2 > 0) {
        r_ = y;
        theta = 0.5 - 0.25*x/y;
    }
    if (ri != 0.0) {
        double q = ri/r;
        double v = sqrt(ri*ri + r_*r_*(1 - q*q));
        r_ = (r_ < 0) ? -v : v;
    }
    // theta is a rational multiple of pi by construction: sincospi needs no argument reduction
    if (!polar) {
        double sn, cs;
        sincospi(theta, &sn, &cs);
        o1 = r_*cs;
        o2 = r_*sn;
    } else {
        if (r_ < 0) theta -= 1.0;
        o1 = fabs(r_);
        o2 = theta;            // in units of pi
    }
}

// Inverse-CDF lookups use a host-built guide table G (one entry per table entry: the bracket at the k-th
// equidistant CDF level) instead of a binary search: one dependent load + a short walk instead of ~14-21
// dependent loads on the D65 (10 000 entries) or image-pixel (up to 2e6 entries) tables.
// Table layout in the generator aux buffer: x[n], F[n], G[n].

// continuous inverse CDF with linear interpolation (random.py:143-157; scipy interp1d kind="linear")
__device__ inline double icdf_linear(const double* __restrict__ x, int n, double X)
{
    const double* __restrict__ F = x + n;
    const double* __restrict__ G = F + n;
    const double F0 = F[0], Fl = F[n - 1];
    int k = (int)((X - F0)/(Fl - F0)*(double)n);
    k = k < 0 ? 0 : (k > n - 1 ? n - 1 : k);
    int lo = (int)G[k];
    while (lo < n - 2 && X >= F[lo + 1]) ++lo;
    while (lo > 0 && X < F[lo]) --lo;
    double dF = F[lo + 1] - F[lo];
    if (!(dF > 0)) return x[lo];
    return x[lo] + (X - F[lo])/dF*(x[lo + 1] - x[lo]);
}

// discrete inverse CDF (random.py:129-140; interp1d kind="next"): first index with F[i] >= X
__device__ inline int icdf_next(const double* __restrict__ x, int n, double X)
{
    const double* __restrict__ F = x + n;
    const double* __restrict__ G = F + n;
    int k = (int)(X/F[n - 1]*(double)n);
    k = k < 0 ? 0 : (k > n - 1 ? n - 1 : k);
    int i = (int)G[k];
    while (i < n - 1 && F[i] < X) ++i;
    while (i > 0 && F[i - 1] >= X) --i;
    return i;
}

#define OTB_GEN_MAXSRC 16
struct GenArgs {
    OtbSource src[OTB_GEN_MAXSRC];     // by value: the source records ride in the constant bank
    int nsrc, no_pol, src_index0, pad;
    const double* aux;
    int64_t N, k_begin, k_end, ray_offset;
    uint64_t seed;
    double *p0, *s0;
    float *pol0, *w0, *wl0;
    int* status;
};

__global__ void __launch_bounds__(128)
generate_kernel(const __grid_constant__ GenArgs a)
{
    const int nsrc = a.nsrc, no_pol = a.no_pol;
    const double* __restrict__ aux = a.aux;
    const int64_t N = a.N, ray_offset = a.ray_offset;
    const uint64_t seed = a.seed;
    double* __restrict__ p0 = a.p0;
    double* __restrict__ s0 = a.s0;
    float* __restrict__ pol0 = a.pol0;
    float* __restrict__ w0 = a.w0;
    float* __restrict__ wl0 = a.wl0;
    int* status = a.status;
    GenCtx g;
    int si_prev = -1;
    for (int64_t k = a.k_begin + (int64_t)blockIdx.x*blockDim.x + threadIdx.x; k < a.k_end; k += (int64_t)gridDim.x*blockDim.x) {
        int si = 0;
        for (int j = 1; j < nsrc; ++j) if (k >= a.src[j].ray_start) si = j;
        const OtbSource& S = a.src[si];
        if (si != si_prev) {           // per-source constants
            si_prev = si;
            const uint64_t n = (uint64_t)S.n_rays;
            feistel_setup(n, g.fa, g.fb);
            uint64_t N2 = (uint64_t)sqrt((double)n);
            while (N2*N2 > n) --N2;
            while ((N2 + 1)*(N2 + 1) <= n) ++N2;
            g.N2 = N2;
            g.inv_n = 1.0/(double)n;
        }
        g.seed = seed;
        g.gid = (uint64_t)(ray_offset + k);
        g.m = (uint64_t)(k - S.ray_start);
        g.n = (uint64_t)S.n_rays;
        g.src = (uint32_t)(a.src_index0 + si);
        g.A = philox4x32_10(g.gid, 100u, g.src, g.seed);
        if (S.shape > OTB_SHAPE_RECT) g.B = philox4x32_10(g.gid, 101u, g.src, g.seed);      // image sources only

        // ---- position (circular_surface.py:32-43, ring_surface.py:135-148, rectangular_surface.py:144-159,
        //      line.py:81-96, point.py:62-69, image sources ray_source.py:237-255)
        double px = S.pos[0], py = S.pos[1];
        const double pz = S.pos[2];
        int pix = -1;
        switch (S.shape) {
        case OTB_SHAPE_POINT: break;
        case OTB_SHAPE_LINE: {
            double t = strat1(g, ST_POS, -S.geom[0], S.geom[0]);
            px = S.pos[0] + S.geom[1]*t;
            py = S.pos[1] + S.geom[2]*t;
            break;
        }
        case OTB_SHAPE_CIRCLE:
        case OTB_SHAPE_RING: {
            double x, y;
            strat_ring(g, ST_POS, S.geom[0], S.geom[1], false, x, y);
            px += x;
            py += y;
            break;
        }
        case OTB_SHAPE_RECT: {
            double x, y;
            strat2(g, ST_POS, -S.geom[0]/2, S.geom[0]/2, -S.geom[1]/2, S.geom[1]/2, x, y);
            if (S.geom[4] != 0.0) {
                double xr = x*S.geom[2] - y*S.geom[3], yr = x*S.geom[3] + y*S.geom[2];
                x = xr;
                y = yr;
            }
            px += x;
            py += y;
            break;
        }
        default: {   // image sources: pixel by discrete inverse CDF of pixel power, uniform offset inside the pixel
            if (S.img_w*S.img_h > 1) {
                const double* idx = aux + S.pix_cdf_off;
                const double* F = idx + S.pix_cdf_n;
                double X = strat1(g, ST_PIX, 0.0, F[S.pix_cdf_n - 1]);
                pix = (int)idx[icdf_next(idx, S.pix_cdf_n, X)];
            } else {
                pix = 0;
            }
            int PY = pix/S.img_w, PX = pix - PY*S.img_w;
            double rx, ry;
            strat2(g, ST_PIXOFF, 0.0, 1.0, 0.0, 1.0, rx, ry);
            px = (S.extent[1] - S.extent[0])/(double)S.img_w*((double)PX + rx) + S.extent[0];
            py = (S.extent[3] - S.extent[2])/(double)S.img_h*((double)PY + ry) + S.extent[2];
            break;
        }
        }

        // ---- wavelength (light_spectrum.py:81-138, srgb.py:513-553)
        double wl;
        switch (S.wl_mode) {
        case OTB_WL_MONO: wl = (double)(float)S.wl[0]; break;
        case OTB_WL_UNIFORM: wl = strat1(g, ST_WL, S.wl[0], S.wl[1]); break;
        case OTB_WL_DISCRETE: {
            const double* x = aux + S.wl_tab_off;
            const double* F = x + S.wl_tab_n;
            wl = x[icdf_next(x, S.wl_tab_n, strat1(g, ST_WL, 0.0, F[S.wl_tab_n - 1]))];
            break;
        }
        case OTB_WL_CDF: {
            const double* x = aux + S.wl_tab_off;
            const double* F = x + S.wl_tab_n;
            wl = icdf_linear(x, S.wl_tab_n, strat1(g, ST_WL, F[0], F[S.wl_tab_n - 1]));
            break;
        }
        case OTB_WL_GAUSSIAN: {
            double X = strat1(g, ST_WL, S.wl[2], S.wl[3]);
            wl = S.wl[0] + 1.4142135623730951*S.wl[1]*erfinv(2*X - 1);
            break;
        }
        default: {   // OTB_WL_SRGB: choose a primary by the pixel's linear-RGB mixing ratios, then its inverse CDF
            const double* th = aux + S.pix_rgb_off + 2*(int64_t)pix;
            double c = strat1(g, ST_RGB, 0.0, 1.0);
            int prim = (c < th[0]) ? 0 : ((c > th[1]) ? 2 : 1);
            const double* x = aux + S.srgb_off + 15000*prim;      // per primary: wl[5000], F[5000], G[5000]
            const double* F = x + 5000;
            wl = icdf_linear(x, 5000, strat1(g, ST_WL, F[0], F[4999]));
            break;
        }
        }

        // ---- orientation (ray_source.py:264-277)
        V3 so;
        if (S.orientation == OTB_OR_CONSTANT) so = v3(S.s[0], S.s[1], S.s[2]);
        else so = unit3(v3(S.conv_pos[0] - px, S.conv_pos[1] - py, S.conv_pos[2] - pz));

        // ---- divergence (ray_source.py:290-351)
        V3 s = so;
        if (S.divergence != OTB_DIV_NONE) {
            // the direction needs sin/cos of theta only: where the sampling law gives them in closed form
            // (asin / acos of the sampled radius) they are computed algebraically instead of through
            // inverse + forward trigonometry; alpha goes through one sincos
            double theta = 0.0, alpha, ct = 0.0, stt = 0.0;
            bool have_sc = false, alpha_pi = false;       // alpha_pi: alpha is given in units of pi
            if (S.div_2d) {
                Philox4 r = draw(g, ST_DIV2);
                // two equally likely half-planes; stratified over the rays like the reference's discrete draw
                alpha = S.div_axis + ((stratum(g, ST_DIV2) & 1) ? 3.141592653589793 : 0.0);
                (void)r;
                if (S.divergence == OTB_DIV_LAMBERTIAN) theta = asin(strat1(g, ST_DIV, 0.0, S.div_sin));
                else if (S.divergence == OTB_DIV_ISOTROPIC) theta = strat1(g, ST_DIV, 0.0, S.div_angle);
                else {
                    const double* x = aux + S.div_tab_off;
                    const double* F = x + S.div_tab_n;
                    theta = icdf_linear(x, S.div_tab_n, strat1(g, ST_DIV, F[0], F[S.div_tab_n - 1]));
                }
            } else {
                double rr;
                strat_ring(g, ST_DIV, 0.0, S.div_sin, true, rr, alpha);
                alpha_pi = true;
                if (S.divergence == OTB_DIV_LAMBERTIAN) {            // theta = asin(r)
                    stt = rr;
                    ct = sqrt(1 - rr*rr);
                    have_sc = true;
                } else if (S.divergence == OTB_DIV_ISOTROPIC) {       // theta = acos(1 - r^2)
                    ct = 1 - rr*rr;
                    stt = rr*sqrt(2 - rr*rr);
                    have_sc = true;
                } else {
                    const double* x = aux + S.div_tab_off;
                    const double* F = x + S.div_tab_n;
                    double X0 = rr*rr/(S.div_sin*S.div_sin);
                    theta = icdf_linear(x, S.div_tab_n, F[0] + X0*(F[S.div_tab_n - 1] - F[0]));
                }
            }
            double fa = 1/sqrt(1 - so.x*so.x);
            V3 sy = v3(0.0, -so.z*fa, so.y*fa);
            V3 sx = cross3(so, sy);
            double ca, sa;
            if (!have_sc) sincos(theta, &stt, &ct);
            if (alpha_pi) sincospi(alpha, &sa, &ca); else sincos(alpha, &sa, &ca);
            s = v3(ct*so.x + stt*(ca*sx.x + sa*sy.x), ct*so.y + stt*(ca*sx.y + sa*sy.y), ct*so.z + stt*(ca*sx.z + sa*sy.z));
        }
        if (!(s.z > 0)) atomicOr(status, OTB_STATUS_NEG_DIR);

        // ---- polarisation (ray_source.py:359-433)
        if (!no_pol) {
            double ang;
            bool ang_pi = false;                           // angle given in units of pi
            switch (S.polarization) {
            case OTB_POL_CONSTANT: ang = S.pol_angle; break;
            case OTB_POL_UNIFORM: ang = strat1(g, ST_POL, 0.0, 2.0); ang_pi = true; break;
            case OTB_POL_LIST: {
                const double* x = aux + S.pol_tab_off;
                const double* F = x + S.pol_tab_n;
                ang = x[icdf_next(x, S.pol_tab_n, strat1(g, ST_POL, 0.0, F[S.pol_tab_n - 1]))];
                break;
            }
            default: {
                const double* x = aux + S.pol_tab_off;
                const double* F = x + S.pol_tab_n;
                ang = icdf_linear(x, S.pol_tab_n, strat1(g, ST_POL, F[0], F[S.pol_tab_n - 1]));
                ang = ang*0.017453292519943295;   // sic: the reference applies np.radians to the sampled angle (ray_source.py:392)
                break;
            }
            }
            double sang, cang;
            if (ang_pi) sincospi(ang, &sang, &cang); else sincos(ang, &sang, &cang);
            V3 pol = v3(cang, sang, 0.0);
            if (s.z != 1) {
                double fa = 1/(sqrt(1 - s.z*s.z) + 1e-16);
                V3 ps = v3(s.y*fa, -s.x*fa, 0.0);
                double A_ts = ps.x*pol.x + ps.y*pol.y;
                double A_tp = ps.y*pol.x - ps.x*pol.y;
                V3 pp_ = cross3(ps, s);
                pol = v3(ps.x*A_ts + pp_.x*A_tp, ps.y*A_ts + pp_.y*A_tp, ps.z*A_ts + pp_.z*A_tp);
            }
            pol0[k] = (float)pol.x;
            pol0[k + N] = (float)pol.y;
            pol0[k + 2*N] = (float)pol.z;
        }

        p0[k] = px;
        p0[k + N] = py;
        p0[k + 2*N] = pz;
        s0[k] = s.x;
        s0[k + N] = s.y;
        s0[k + 2*N] = s.z;
        w0[k] = (float)S.weight;
        wl0[k] = (float)wl;
    }
}

int otb_sm_count();

extern "C" int otb_generate_rays(const OtbSource* sources_h, int n_sources, const double* gen_aux_d, int64_t N,
                                 uint64_t seed, int64_t ray_offset, int no_pol, double* p0_d, double* s0_d,
                                 float* pol0_d, float* w0_d, float* wl_d, int32_t* status_d, void* stream)
{
    if (!sources_h || n_sources < 1 || !p0_d || !s0_d || !w0_d || !wl_d || (!no_pol && !pol0_d) || !status_d) {
        otb_set_error("null argument");
        return OTB_ERR_INVALID_ARG;
    }
    if (N <= 0) return OTB_OK;
    int64_t cover = 0;
    for (int i = 0; i < n_sources; ++i) {
        if (sources_h[i].ray_start != cover || sources_h[i].n_rays < 0) {
            otb_set_error("sources must cover [0, N) with contiguous blocks (RayStorage.B_list)");
            return OTB_ERR_INVALID_ARG;
        }
        cover += sources_h[i].n_rays;
    }
    if (cover != N) { otb_set_error("source ray counts do not sum to N"); return OTB_ERR_INVALID_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    // one launch per group of <= 16 sources over the ray range that group covers; no allocation, no sync
    for (int g0 = 0; g0 < n_sources; g0 += OTB_GEN_MAXSRC) {
        GenArgs a;
        a.nsrc = (n_sources - g0 < OTB_GEN_MAXSRC) ? n_sources - g0 : OTB_GEN_MAXSRC;
        for (int i = 0; i < a.nsrc; ++i) a.src[i] = sources_h[g0 + i];
        a.no_pol = no_pol;
        a.src_index0 = g0;
        a.aux = gen_aux_d;
        a.N = N;
        a.k_begin = sources_h[g0].ray_start;
        a.k_end = sources_h[g0 + a.nsrc - 1].ray_start + sources_h[g0 + a.nsrc - 1].n_rays;
        a.ray_offset = ray_offset;
        a.seed = seed;
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
