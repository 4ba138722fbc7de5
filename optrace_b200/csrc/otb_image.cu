// otb_image.cu — the step right after the detector histogram: RenderImage.get (render_image.py:131-222) on
// the device.  Join-bins rescaling of the (Ny, Nx, 4) XYZW histogram (cv2.resize INTER_AREA with an integer
// factor = block mean) and the per-pixel conversions to irradiance / illuminance / sRGB (absolute and perceptual
// rendering intent) / CIELUV lightness, hue, chroma, saturation / out-of-gamut mask, following
// color/srgb.py:118-407, color/luv.py:20-139 and color/xyz.py:17-35 operation by operation in fp64.
//
// Conversions that normalise by an image-wide extremum run as: one statistics pass (block reduction + one atomic
// per block and statistic), the host derives the scalars the reference derives, one conversion pass.  The
// perceptual intent needs two more statistics that depend on the first ones (srgb.py:238-262, 331-352).
// All passes are HBM-bound streaming kernels over at most 4725 x 945 pixels; no tensor cores (no contraction).
#include "otb_common.cuh"

// ---- colour constants (srgb.py:19-21, xyz.py:10-13, luv.py:8-17) -----------------------------------------
#define WP_X 0.31272
#define WP_Y 0.32903
#define WP_UN 0.19783982
#define WP_VN 0.4683363

struct XYZ3 { double X, Y, Z; };

// _to_srgb without normalisation (srgb.py:118-136)
__device__ __forceinline__ void to_rgbl(double X, double Y, double Z, double& r, double& g, double& b)
{
    r = 3.2404542*X + -1.5371385*Y + -0.4985314*Z;
    g = -0.9692660*X + 1.8760108*Y + 0.0415560*Z;
    b = 0.0556434*X + -0.2040259*Y + 1.0572252*Z;
}

// _triangle_intersect (srgb.py:139-187): project (x, y) towards the white point w onto the gamut triangle r, g, b
__device__ __forceinline__ void triangle_intersect(double rx, double ry, double gx, double gy, double bx, double by,
                                                   double wx, double wy, double& x, double& y)
{
    const double phir = atan2(ry - wy, rx - wx);
    const double phig = atan2(gy - wy, gx - wx);
    const double phib = atan2(by - wy, bx - wx) + 2*3.141592653589793;
    double phi = atan2(y - wy, x - wx);
    if (phi < 0) phi += 2*3.141592653589793;
    const double aw = tan(phi);
    const double abg = (gy - by)/(gx - bx);
    const double abr = (ry - by)/(rx - bx);
    const double agr = (ry - gy)/(rx - gx);
    const bool is_bg = (phi <= phib) && (phi > phig);
    const bool is_gr = (phi <= phig) && (phi > phir);
    if (is_bg) {
        x = (y - x*aw + (bx*abg - by))/(abg - aw);
        y = x*abg + (by - bx*abg);
    } else if (is_gr) {
        x = (y - x*aw + (gx*agr - gy))/(agr - aw);
        y = x*agr + (gy - gx*agr);
    } else {
        x = (y - x*aw + (bx*abr - by))/(abr - aw);
        y = x*abr + (by - bx*abr);
    }
}

// saturation clipping of one out-of-gamut colour (srgb.py:313-329): hue and Y stay, chroma is reduced
__device__ __forceinline__ void absolute_fix(double& X, double Y, double& Z)
{
    const double s = X + Y + Z;                                   // xyz_to_xyY (xyz.py:17-35)
    double x = (s > 0) ? X/s : WP_X, y = (s > 0) ? Y/s : WP_Y;
    triangle_intersect(0.64, 0.33, 0.30, 0.60, 0.15, 0.06, WP_X, WP_Y, x, y);
    const double k = Y/((y > 0) ? y : INFINITY);
    X = k*x;
    Z = k*(1 - x - y);
}

// xyz_to_luv (luv.py:20-69) for one pixel, Yn = normalisation luminance
__device__ __forceinline__ void xyz_to_luv(double X, double Y, double Z, double Yn, double& L, double& u, double& v)
{
    X = fmax(X, 0.0); Y = fmax(Y, 0.0); Z = fmax(Z, 0.0);
    L = u = v = 0.0;
    if (!(Y > 0)) return;
    const double t = 1/Yn*Y;
    L = (t > 0.008856) ? 116*pow(t, 1.0/3) - 16 : 903.3*t;
    const double D = 1/(X + 15*Y + 3*Z);
    const double uu = 4*X*D, vv = 9*Y*D;
    const double L13 = 13*L;
    u = L13*(uu - WP_UN);
    v = L13*(vv - WP_VN);
}

// luv_to_xyz (luv.py:72-106)
__device__ __forceinline__ void luv_to_xyz(double L, double u, double v, double& X, double& Y, double& Z)
{
    X = Y = Z = 0.0;
    if (!(L > 0)) return;
    Y = (L > 903.3*0.008856) ? pow(1.0/116*(L + 16), 3.0) : 1/903.3*L;
    const double L13 = 13*L;
    X = 9.0/4*Y*(u + L13*WP_UN)/(v + L13*WP_VN);
    Z = 3*Y*(L13/(v + L13*WP_VN) - 5.0/3) - 1.0/3*X;
}

// luv_to_u_v_l (luv.py:109-124) + _get_chroma_scale (srgb.py:190-235) for one pixel: visible-gamut test and the
// squared chroma scaling factor that brings the colour onto the sRGB triangle in the u'v' diagram
__device__ __forceinline__ void chroma_scale_px(double L, double u, double v, bool& in_gamut, double& cr_fact2)
{
    double u_ = WP_UN, v_ = WP_VN;
    if (L > 0) {
        u_ += 1.0/13*u/L;
        v_ += 1.0/13*v/L;
    }
    const bool l1 = v_ > (0.5065 - 0.013)/(0.6235 - 0.255)*(u_ - 0.2555) + 0.01373;
    const bool l2 = v_ < (0.5065 - 0.6)/(0.6235 - 0.0)*u_ + 0.6;
    const bool l3 = u_ > 0;
    const bool l4 = v_ > (0.013 - 0.28)/(0.255 - 0)*u_ + 0.28;
    const bool l5 = v_ > (0.0 - 0.48)/(0.18 - 0)*u_ + 0.48;
    in_gamut = l1 && l2 && l3 && l4 && l5;
    const double cr0 = (u_ - WP_UN)*(u_ - WP_UN) + (v_ - WP_VN)*(v_ - WP_VN);
    triangle_intersect(0.4507042254, 0.5228873239, 0.125, 0.5625, 0.1754385965, 0.1578947368, WP_UN, WP_VN, u_, v_);
    const double cr1 = (u_ - WP_UN)*(u_ - WP_UN) + (v_ - WP_VN)*(v_ - WP_VN);
    cr_fact2 = cr1/(cr0 + 1e-9);
}

// srgb_linear_to_srgb (srgb.py:355-375) after np.clip(RGBL, 0, 1) (srgb.py:402-403)
__device__ __forceinline__ double gamma_srgb(double c)
{
    c = fmin(fmax(c, 0.0), 1.0);
    if (fabs(c) <= 0.0031308) return c*12.92;
    return (1 + 0.055)*pow(fabs(c), 1/2.4) - 0.055;
}

// ---- block reductions into a statistics record --------------------------------------------------------------
__device__ __forceinline__ void atomic_max_f64(double* addr, double v)
{
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (!(v > __longlong_as_double((long long)assumed))) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_min_f64(double* addr, double v)
{
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (!(v < __longlong_as_double((long long)assumed))) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

template <int NS>
__device__ __forceinline__ void block_reduce_store(double (&v)[NS], const bool (&is_min)[NS], double* stats)
{
    __shared__ double sm[NS][8];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, v[k], d);
            v[k] = is_min[k] ? fmin(v[k], o) : fmax(v[k], o);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < NS; ++k) sm[k][warp] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < NS; ++k) {
            double r = sm[k][0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_min[k] ? fmin(r, sm[k][w]) : fmax(r, sm[k][w]);
            if (is_min[k]) atomic_min_f64(&stats[k], r); else atomic_max_f64(&stats[k], r);
        }
    }
}

// statistics record (doubles):
//  [0] max of linear RGB, untouched colours            (nanmax in _to_srgb, srgb.py:133)
//  [1] 1 when some colour lies outside the sRGB gamut  (np.any(inv), srgb.py:303-306)
//  [2] max of Y over pixels with Y > 0                 (Yn of xyz_to_luv(normalize=True), luv.py:40)
//  [3] max of L with Yn = 1                            (Luv[:, :, 0].max(), srgb.py:256)
//  [4] 1 when some colour lies inside the visible gamut test (np.any(in_gamut), srgb.py:220)
//  [5] max of linear RGB after the absolute-intent fix (second _to_srgb, srgb.py:354)
//  [6] min of cr_fact2 over valid colours above the lightness threshold, +inf when none (srgb.py:257-259)
//  [7] max of linear RGB after the perceptual chroma scaling
#define OTB_IMG_NSTATS 8

__global__ void __launch_bounds__(256) image_stats1_kernel(const double* __restrict__ img, int64_t npx, double* stats)
{
    double v[6] = {-INFINITY, 0.0, -INFINITY, -INFINITY, 0.0, -INFINITY};
    const bool is_min[6] = {false, false, false, false, false, false};
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x*blockDim.x) {
        const double X = img[4*i], Y = img[4*i + 1], Z = img[4*i + 2];
        double r, g, b;
        to_rgbl(X, Y, Z, r, g, b);
        v[0] = fmax(v[0], fmax(r, fmax(g, b)));
        const bool inv = (r < 0) || (g < 0) || (b < 0);
        if (inv) v[1] = 1.0;
        const double Yc = fmax(Y, 0.0);
        if (Yc > 0) v[2] = fmax(v[2], Yc);
        double L, u, w;
        xyz_to_luv(X, Y, Z, 1.0, L, u, w);
        v[3] = fmax(v[3], L);
        bool ing;
        double cf2;
        chroma_scale_px(L, u, w, ing, cf2);
        if (ing) v[4] = 1.0;
        double X2 = X, Z2 = Z;
        if (inv) absolute_fix(X2, Y, Z2);
        to_rgbl(X2, Y, Z2, r, g, b);
        v[5] = fmax(v[5], fmax(r, fmax(g, b)));
    }
    block_reduce_store<6>(v, is_min, stats);
}

// perceptual intent, second statistics: needs max L (stats[3]) and the threshold
__global__ void __launch_bounds__(256) image_stats2_kernel(const double* __restrict__ img, int64_t npx, double L_th, double* stats)
{
    double v[1] = {INFINITY};
    const bool is_min[1] = {true};
    const double Lmin = L_th*stats[3];
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x*blockDim.x) {
        double L, u, w;
        xyz_to_luv(img[4*i], img[4*i + 1], img[4*i + 2], 1.0, L, u, w);
        bool ing;
        double cf2;
        chroma_scale_px(L, u, w, ing, cf2);
        if (ing && L > Lmin) v[0] = fmin(v[0], cf2);
    }
    block_reduce_store<1>(v, is_min, stats + 6);
}

// perceptual intent: chroma-scaled colour of one pixel (srgb.py:331-352)
__device__ __forceinline__ void perceptual_px(double X, double Y, double Z, double chroma_scale, bool any_gamut,
                                              double& r, double& g, double& b)
{
    double L, u, v;
    xyz_to_luv(X, Y, Z, 1.0, L, u, v);
    bool ing;
    double cf2 = 1.0;
    if (any_gamut) chroma_scale_px(L, u, v, ing, cf2);
    double cf = sqrt(cf2);
    if (cf > chroma_scale) cf = chroma_scale;
    u *= cf;
    v *= cf;
    double X2, Y2, Z2;
    luv_to_xyz(L, u, v, X2, Y2, Z2);
    to_rgbl(X2, Y2, Z2, r, g, b);
}

__global__ void __launch_bounds__(256) image_stats3_kernel(const double* __restrict__ img, int64_t npx, double chroma_scale,
                                                           double* stats)
{
    double v[1] = {-INFINITY};
    const bool is_min[1] = {false};
    const bool any_gamut = stats[4] != 0.0;
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x*blockDim.x) {
        double r, g, b;
        perceptual_px(img[4*i], img[4*i + 1], img[4*i + 2], chroma_scale, any_gamut, r, g, b);
        v[0] = fmax(v[0], fmax(r, fmax(g, b)));
    }
    block_reduce_store<1>(v, is_min, stats + 7);
}

struct ConvertArgs {
    const double* img;      // (H, W, 4)
    double* out;            // (H, W) or (H, W, 3)
    const double* stats;
    int64_t npx;
    int mode;
    double scale;           // Irradiance: 1/Apx, Illuminance: K/Apx
    double chroma_scale;    // perceptual intent; < 0: plain conversion (no colour outside the gamut)
};

__global__ void __launch_bounds__(256) image_convert_kernel(const ConvertArgs a)
{
    const double* __restrict__ img = a.img;
    const double* __restrict__ st = a.stats;
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < a.npx; i += (int64_t)gridDim.x*blockDim.x) {
        const double X = img[4*i], Y = img[4*i + 1], Z = img[4*i + 2], W = img[4*i + 3];
        switch (a.mode) {
        case OTB_IMG_IRRADIANCE: a.out[i] = a.scale*W; break;
        case OTB_IMG_ILLUMINANCE: a.out[i] = a.scale*Y; break;
        case OTB_IMG_SRGB_ABSOLUTE: {
            double r, g, b, X2 = X, Z2 = Z;
            to_rgbl(X, Y, Z, r, g, b);
            if ((r < 0) || (g < 0) || (b < 0)) absolute_fix(X2, Y, Z2);
            to_rgbl(X2, Y, Z2, r, g, b);
            const double nmax = st[5];
            if (nmax != 0.0 && nmax == nmax) {       // `if normalize and (nmax := np.nanmax(RGBL_))`
                const double f = 1/nmax;
                r *= f; g *= f; b *= f;
            }
            a.out[3*i] = gamma_srgb(r);
            a.out[3*i + 1] = gamma_srgb(g);
            a.out[3*i + 2] = gamma_srgb(b);
            break;
        }
        case OTB_IMG_SRGB_PERCEPTUAL: {
            double r, g, b, nmax;
            if (a.chroma_scale < 0) {                // srgb.py:308-310: everything in gamut, nothing to scale
                to_rgbl(X, Y, Z, r, g, b);
                nmax = st[0];
            } else {
                perceptual_px(X, Y, Z, a.chroma_scale, st[4] != 0.0, r, g, b);
                nmax = st[7];
            }
            if (nmax != 0.0 && nmax == nmax) {
                const double f = 1/nmax;
                r *= f; g *= f; b *= f;
            }
            a.out[3*i] = gamma_srgb(r);
            a.out[3*i + 1] = gamma_srgb(g);
            a.out[3*i + 2] = gamma_srgb(b);
            break;
        }
        case OTB_IMG_OUTSIDE_GAMUT: {
            double r, g, b;
            to_rgbl(X, Y, Z, r, g, b);
            const double nmax = st[0];
            if (nmax != 0.0 && nmax == nmax) {
                const double f = 1/nmax;
                r *= f; g *= f; b *= f;
            }
            a.out[i] = ((r < -1e-6) || (g < -1e-6) || (b < -1e-6)) ? 1.0 : 0.0;
            break;
        }
        default: {      // CIELUV family: normalised by the largest Y (luv.py:39-42); all zero when no pixel has Y > 0
            double L = 0.0, u = 0.0, v = 0.0;
            if (st[2] > 0) xyz_to_luv(X, Y, Z, st[2], L, u, v);
            double o;
            if (a.mode == OTB_IMG_LIGHTNESS) o = L;
            else if (a.mode == OTB_IMG_HUE) {
                o = 180/3.141592653589793*atan2(v, u);
                if (o < 0) o += 360;
            } else {
                const double C = sqrt(u*u + v*v);
                o = (a.mode == OTB_IMG_CHROMA) ? C : ((L > 0) ? C/L : 0.0);
            }
            a.out[i] = o;
            break;
        }
        }
    }
}

// join bins: every output pixel is the mean of a fact x fact block (cv2.resize INTER_AREA, render_image.py:170-174)
__global__ void __launch_bounds__(256) image_rescale_kernel(const double* __restrict__ img, int Ny, int Nx, int fact, double* __restrict__ out)
{
    const int Wo = Nx/fact, Ho = Ny/fact;
    const int64_t n = (int64_t)Wo*Ho*4;
    // OpenCV multiplies the bin sum by `float scale = 1.f/(fact*fact)` (resizeAreaFast_): the float32 rounding of
    // the scale is part of the reference's numbers (1e-8 relative in irradiance / illuminance)
    const double inv = (double)(1.0f/(float)(fact*fact));
    for (int64_t k = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x*blockDim.x) {
        const int c = (int)(k & 3);
        const int64_t px = k >> 2;
        const int xo = (int)(px % Wo), yo = (int)(px/Wo);
        double s = 0.0;
        for (int dy = 0; dy < fact; ++dy) {
            const double* row = img + (((int64_t)(yo*fact + dy))*Nx + (int64_t)xo*fact)*4 + c;
            for (int dx = 0; dx < fact; ++dx) s += row[4*dx];
        }
        out[k] = s*inv;
    }
}

// Resolution-limit filter (RenderImage._apply_rayleigh_filter, render_image.py:255-296): the XYZW histogram is
// convolved with the Airy-disc kernel the host built (scipy.special.j1, a (2 ps + 1)^2 table), zero padded
// ("same"), negatives removed.  The reference goes through an FFT; the kernel has compact support (third zero of
// the Airy pattern), so this is a direct convolution: thread = pixel (4 channels = two 128-bit loads per tap),
// taps outside the support are skipped warp-uniformly, the PSF table is read through the read-only cache.
__global__ void __launch_bounds__(256) image_convolve_kernel(const double* __restrict__ img, int Ny, int Nx,
                                                             const double* __restrict__ psf, int K, double* __restrict__ out)
{
    const int x = blockIdx.x*32 + (threadIdx.x & 31), y = blockIdx.y*8 + (threadIdx.x >> 5);
    if (x >= Nx || y >= Ny) return;
    const int ps = K >> 1;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int j = 0; j < K; ++j) {
        const int yy = y + ps - j;                  // out[y] = sum_j in[y + ps - j] * psf[j]  (true convolution)
        if (yy < 0 || yy >= Ny) continue;
        const double* __restrict__ prow = psf + (int64_t)j*K;
        const double2* __restrict__ irow = (const double2*)(img + (int64_t)yy*Nx*4);
        for (int i = 0; i < K; ++i) {
            const double pv = __ldg(prow + i);
            if (pv == 0.0) continue;
            const int xx = x + ps - i;
            if (xx < 0 || xx >= Nx) continue;
            const double2 v01 = __ldg(irow + 2*xx), v23 = __ldg(irow + 2*xx + 1);
            a0 += v01.x*pv;
            a1 += v01.y*pv;
            a2 += v23.x*pv;
            a3 += v23.y*pv;
        }
    }
    double* o = out + ((int64_t)y*Nx + x)*4;        // "remove negative values that can arise by fft": none arise here
    o[0] = a0 < 0 ? 0.0 : a0;
    o[1] = a1 < 0 ? 0.0 : a1;
    o[2] = a2 < 0 ? 0.0 : a2;
    o[3] = a3 < 0 ? 0.0 : a3;
}

int otb_sm_count();

extern "C" {

int otb_image_rescale(const double* img_d, int32_t Ny, int32_t Nx, int32_t fact, double* out_d, void* stream)
{
    if (!img_d || !out_d) { otb_set_error("null argument"); return OTB_ERR_INVALID_ARG; }
    if (fact < 1 || Ny < 1 || Nx < 1 || Ny % fact || Nx % fact) { otb_set_error("rescale factor must divide both image sides"); return OTB_ERR_INVALID_ARG; }
    const int64_t n = (int64_t)(Nx/fact)*(Ny/fact)*4;
    const int blocks = (int)((n + 255)/256 < 16LL*otb_sm_count() ? (n + 255)/256 : 16LL*otb_sm_count());
    image_rescale_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(img_d, Ny, Nx, fact, out_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

// pass: 1 = first statistics ([0..5], resets the whole record), 2 = [6] (needs L_th), 3 = [7] (needs chroma_scale)
int otb_image_stats(const double* img_d, int64_t npx, int32_t pass, double param, double* stats_d, void* stream)
{
    if (!img_d || !stats_d || npx < 1) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)((npx + 255)/256 < 8LL*otb_sm_count() ? (npx + 255)/256 : 8LL*otb_sm_count());
    if (pass == 1) {
        const double init[OTB_IMG_NSTATS] = {-INFINITY, 0.0, -INFINITY, -INFINITY, 0.0, -INFINITY, INFINITY, -INFINITY};
        OTB_CUDA(cudaMemcpyAsync(stats_d, init, sizeof(init), cudaMemcpyHostToDevice, st));
        image_stats1_kernel<<<blocks, 256, 0, st>>>(img_d, npx, stats_d);
    } else if (pass == 2) {
        image_stats2_kernel<<<blocks, 256, 0, st>>>(img_d, npx, param, stats_d);
    } else if (pass == 3) {
        image_stats3_kernel<<<blocks, 256, 0, st>>>(img_d, npx, param, stats_d);
    } else {
        otb_set_error("invalid statistics pass");
        return OTB_ERR_INVALID_ARG;
    }
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

int otb_image_convert(const double* img_d, int64_t npx, int32_t mode, double scale, double chroma_scale,
                      const double* stats_d, double* out_d, void* stream)
{
    if (!img_d || !out_d || !stats_d || npx < 1) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    if (mode < OTB_IMG_IRRADIANCE || mode > OTB_IMG_SATURATION) { otb_set_error("invalid image mode"); return OTB_ERR_INVALID_ARG; }
    ConvertArgs a;
    a.img = img_d; a.out = out_d; a.stats = stats_d; a.npx = npx; a.mode = mode; a.scale = scale; a.chroma_scale = chroma_scale;
    const int blocks = (int)((npx + 255)/256 < 8LL*otb_sm_count() ? (npx + 255)/256 : 8LL*otb_sm_count());
    image_convert_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

int otb_image_convolve(const double* img_d, int32_t Ny, int32_t Nx, const double* psf_d, int32_t K, double* out_d, void* stream)
{
    if (!img_d || !psf_d || !out_d || img_d == out_d) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    if (Ny < 1 || Nx < 1 || K < 1 || !(K & 1)) { otb_set_error("the filter kernel needs an odd side length"); return OTB_ERR_INVALID_ARG; }
    dim3 grid((Nx + 31)/32, (Ny + 7)/8);
    image_convolve_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img_d, Ny, Nx, psf_d, K, out_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

}  // extern "C"
