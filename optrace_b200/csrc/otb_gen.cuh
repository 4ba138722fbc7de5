// otb_gen.cuh — one ray of RaySource.create_rays (ray_source.py:204-437) as a device function: the sampling
// primitives of random.py (stratified grids + shuffle, Shirley disc map, inverse-CDF tables) driven by counter-based
// Philox4x32-10 instead of numpy's SFC64 stream.  Used by the stand-alone generator (otb_gen.cu) and, fused, by
// the prologue of the trace kernels (otb_trace.cu, otb_render.cu): a generated bundle then never touches HBM.
//
// Stratification: the reference draws "grid cell + dither" and then shuffles globally (random.py:25-45).
// Here ray m of a source takes grid cell perm_key(m), a keyed Feistel bijection of [0, n) (otb_rng.cuh), so
// every cell is used exactly once and different random variables are decorrelated by different keys.
// Generated bundles therefore agree with the reference statistically (same distributions, same
// stratification), not bit-wise; parity tests inject reference-generated bundles instead.
#pragma once
#include "otb_common.cuh"
#include "otb_rng.cuh"

#define OTB_GEN_MAXSRC 12
// source records by value: they ride in the constant bank of the kernel that generates (<= 4.2 KB)
// per-source constants of the stratum permutations, computed once on the host (otb_fill_genblock): they depend on
// the ray count of the source only, and as kernel parameters they cost no registers in the trace loop
struct GenSrcConst {
    uint32_t fa, fb;       // moduli of the stratum permutation (feistel_setup)
    uint32_t ga, gb;       // moduli of the block permutation of the coherent variable
    uint64_t N2;           // floor(sqrt(n)): side of the stratified square grid (random.py:25-33)
    double inv_n;          // 1 / n
    uint32_t coh_stream;   // random variable whose strata are handed out in blocks of 32 consecutive cells (0: none)
    uint32_t pad;
};

struct GenBlock {
    OtbSource src[OTB_GEN_MAXSRC];
    GenSrcConst sc[OTB_GEN_MAXSRC];
    int nsrc, no_pol, src_index0, pad;
    const double* aux;
    int64_t ray_offset;
    uint64_t seed;
};

struct GenRay {
    V3 p, s;
    float pol[3];
    float w, wl;
    bool neg_dir;      // direction with s_z <= 0 (ray_source.py:353): the caller raises OTB_STATUS_NEG_DIR
};

struct GenCtx {
    uint64_t seed;
    uint64_t gid;      // global ray id (Philox counter)
    uint64_t m;        // index of the ray inside its source
    uint64_t n;        // rays of this source in this launch
    uint32_t src;      // source index (decorrelates sources)
    // per-source constants (copied from GenSrcConst of the ray's source)
    uint32_t fa, fb;
    uint64_t N2;
    double inv_n;
    Philox4 A, B;      // shared random blocks of this ray (see strat1)
    uint32_t coh_stream;
    uint32_t ga, gb;
};

enum { ST_POS = 1, ST_WL = 2, ST_RGB = 3, ST_DIV = 4, ST_DIV2 = 5, ST_POL = 6, ST_PIX = 7, ST_PIXOFF = 8 };

__device__ __forceinline__ Philox4 draw(const GenCtx& g, uint32_t stream) { return philox4x32_10(g.gid, stream, g.src, g.seed); }
// Stratum of ray m for one random variable.  Default: a keyed bijection of [0, n) per variable (the reference
// shuffles every stratified sample globally, random.py:41-45).  Coherent variable (OtbSource.coherent, see otb.h):
// the bijection acts on BLOCKS of 32 strata, ray m takes cell 32*perm(m / 32) + m % 32 — the 32 rays of a warp
// then sample neighbouring cells of that one variable (all other variables stay fully shuffled against it), so a
// warp's rays share their fate at stops and lens edges and dead warps skip the surface arithmetic.  Every cell is
// still used exactly once: the bundle as a whole has the same stratified distribution.
// out of line on purpose: a ray draws up to eight strata and every inlined copy of the permutation is ~150
// instructions of straight-line code executed once per ray (instruction fetch, not arithmetic, is what it costs)
static __device__ __noinline__ uint64_t stratum_perm(uint64_t m, uint64_t n, uint32_t fa, uint32_t fb, uint32_t ga, uint32_t gb,
                                                    bool coherent, uint64_t key)
{
    if (coherent) {
        const uint64_t nb = n >> 5;                           // full blocks; the tail rays keep their own cells
        const uint64_t blk = m >> 5;
        if (blk < nb) return (feistel_perm_ab(blk, nb, ga, gb, key) << 5) | (m & 31);
        return m;
    }
    return feistel_perm_ab(m, n, fa, fb, key);
}

__device__ __forceinline__ uint64_t stratum(const GenCtx& g, uint32_t stream)
{
    const uint64_t key = g.seed ^ ((uint64_t)stream << 40) ^ ((uint64_t)g.src << 20) ^ 0x5bd1e995u;
    return stratum_perm(g.m, g.n, g.fa, g.fb, g.ga, g.gb, stream == g.coh_stream, key);
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


// per-source constants of the permutations (host side, once per launch)
inline void gen_source_setup(const OtbSource& S, GenSrcConst& g)
{
    const uint64_t n = (uint64_t)S.n_rays;
    feistel_setup(n, g.fa, g.fb);
    uint64_t N2 = (uint64_t)sqrt((double)n);
    while (N2*N2 > n) --N2;
    while ((N2 + 1)*(N2 + 1) <= n) ++N2;
    g.N2 = N2;
    g.inv_n = n ? 1.0/(double)n : 0.0;
    g.coh_stream = 0;
    g.ga = g.gb = 0;
    g.pad = 0;
    if (S.coherent && (n >> 5) > 1) {
        // the variable that decides a ray's fate at stops: the direction inside the divergence cone when there is
        // one, else the position on the source area
        if (S.divergence != OTB_DIV_NONE && !S.div_2d) g.coh_stream = ST_DIV;
        else if (S.shape == OTB_SHAPE_CIRCLE || S.shape == OTB_SHAPE_RING || S.shape == OTB_SHAPE_RECT) g.coh_stream = ST_POS;
        if (g.coh_stream) feistel_setup(n >> 5, g.ga, g.gb);
    }
}

// ray k (local index of this launch) of the sources in G
__device__ __forceinline__ void generate_ray(const GenBlock& G, const int64_t k, GenRay& out)
{
    const double* __restrict__ aux = G.aux;
    int si = 0;
    for (int j = 1; j < G.nsrc; ++j) if (k >= G.src[j].ray_start) si = j;
    const OtbSource& S = G.src[si];
    const GenSrcConst& C = G.sc[si];
    GenCtx g;
    g.fa = C.fa; g.fb = C.fb; g.ga = C.ga; g.gb = C.gb;
    g.N2 = C.N2;
    g.inv_n = C.inv_n;
    g.coh_stream = C.coh_stream;
    g.seed = G.seed;
    g.gid = (uint64_t)(S.gid_start + (k - S.ray_start));       // global ray id: the bundle does not depend on the sharding
    g.m = (uint64_t)(k - S.ray_start);
    g.n = (uint64_t)S.n_rays;
    g.src = (uint32_t)(G.src_index0 + si);
    g.A = philox4x32_10(g.gid, 100u, g.src, g.seed);
    if (S.shape > OTB_SHAPE_RECT) g.B = philox4x32_10(g.gid, 101u, g.src, g.seed);      // image sources only
    out.pol[0] = out.pol[1] = out.pol[2] = 0.0f;

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
    else if (S.orientation == OTB_OR_CONVERGING) so = unit3(v3(S.conv_pos[0] - px, S.conv_pos[1] - py, S.conv_pos[2] - pz));
    else otb_user_v3(S.or_func_id, px, py, &so.x, &so.y, &so.z);        // or_func(x, y) (ray_source.py:274-276)

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
    out.neg_dir = !(s.z > 0);

    // ---- polarisation (ray_source.py:359-433)
    if (!G.no_pol) {
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
        out.pol[0] = (float)pol.x;
        out.pol[1] = (float)pol.y;
        out.pol[2] = (float)pol.z;
    }


    out.p = v3(px, py, pz);
    out.s = s;
    out.w = (float)S.weight;
    out.wl = (float)wl;
}
