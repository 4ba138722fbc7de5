// otb_surfaces.cuh — per-ray surface primitives: mask, height, normal, intersection.
// One thread = one ray; the surface record is warp-uniform, so the kind switches do not diverge.
// References (relative to the reference repository): optrace/tracer/geometry/surface/*.py
#pragma once
#include "otb_common.cuh"

// Capability level of a kernel instantiation: scenes made of flat and conic surfaces only (the usual lens
// systems) run a lean instantiation without the numeric hit finder, splines, user functions and tilted planes:
// fewer registers, a third of the code size (see profiles/).
#define OTB_CAPS_LENS 0     // flat kinds + conic/sphere
#define OTB_CAPS_FULL 1     // + tilted, asphere, function, data surfaces
#define OTB_CAPS_DET 2      // flat kinds + conic + tilted: what a DETECTOR surface can be (detector.py:37-41); fully inlined
                            // code without the out-of-line height functions (calls cost the detector kernels their registers)

struct HitResult {
    V3 p;
    bool hit;
    bool ill;
};

// Surface._rotate_rc (surface.py:427-434) with host-precomputed cos/sin
__device__ __forceinline__ void rot_rc(const KSurface& S, int cslot, double x, double y, double& xr, double& yr)
{
    if (S.flags & OTB_SF_ROTATED) {
        double c = S.par[cslot], s = S.par[cslot + 1];
        xr = x*c - y*s;
        yr = x*s + y*c;
    } else {
        xr = x;
        yr = y;
    }
}

// ------------------------------------------------------------------------------------------------
// quartic B-splines (DataSurface: FITPACK knots/coefficients, data_surface_2d.py:76, 104)
// ------------------------------------------------------------------------------------------------
// index l with t[l] <= x < t[l+1], clamped to [k, n-k-2] (FITPACK splev / fpbisp interval search)
__device__ __forceinline__ int bspl_interval(const double* __restrict__ t, int n, int k, double x)
{
    int lo = k, hi = n - k - 1;          // invariant: answer in [lo, hi-1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x >= t[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// de Boor–Cox recursion (FITPACK fpbspl): the deg+1 non-zero basis functions of degree `deg` (3 or 4) at x, in
// h[0..deg] (h[4] = 0 for deg 3).  Compact on purpose: the level loop is rolled and the five lanes of a level are
// predicated, so the whole recursion is ~150 instructions of code instead of a fully unrolled copy per degree and
// call site — the numeric-surface kernels are bound by instruction fetch (profiles/r2_zoo_numeric_full_ncu.csv), and
// every inlined copy of the surface-kind switch carries this code.  Same operations in the same order as before.
__device__ __forceinline__ void bspl_basis(const double* __restrict__ t, int l, double x, int deg, double* h)
{
    double h0 = 1.0, h1 = 0.0, h2 = 0.0, h3 = 0.0, h4 = 0.0;
#pragma unroll 1
    for (int j = 1; j <= deg; ++j) {
        const double g0 = h0, g1 = h1, g2 = h2, g3 = h3;
        double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0, n4 = 0.0;
        // i = 0 .. j-1:  f = hh[i]/(t[l+i+1] - t[l+i+1-j]);  h[i] += f*(t[l+i+1] - x);  h[i+1] = f*(x - t[l+i+1-j])
        {
            const double ta = t[l + 1], tb = t[l + 1 - j];
            const double f = g0/(ta - tb);
            n0 = n0 + f*(ta - x);
            n1 = f*(x - tb);
        }
        if (j > 1) {
            const double ta = t[l + 2], tb = t[l + 2 - j];
            const double f = g1/(ta - tb);
            n1 = n1 + f*(ta - x);
            n2 = f*(x - tb);
        }
        if (j > 2) {
            const double ta = t[l + 3], tb = t[l + 3 - j];
            const double f = g2/(ta - tb);
            n2 = n2 + f*(ta - x);
            n3 = f*(x - tb);
        }
        if (j > 3) {
            const double ta = t[l + 4], tb = t[l + 4 - j];
            const double f = g3/(ta - tb);
            n3 = n3 + f*(ta - x);
            n4 = f*(x - tb);
        }
        h0 = n0; h1 = n1; h2 = n2; h3 = n3; h4 = n4;
    }
    h[0] = h0; h[1] = h1; h[2] = h2; h[3] = h3; h[4] = h4;
}

// 1-D spline value (nu = 0) or first derivative (nu = 1); extrapolates like splev(ext=0)
// Out of line (like spline2d and surf_values_full below): the numeric-surface kernels are bound by instruction fetch,
// so every table / user-function evaluation exists once per kernel image and is CALLED from the hit finder, the edge
// continuation and the central-difference normals instead of being inlined ~20 times.
static __device__ __noinline__ double spline1d(const double* __restrict__ t, int n, const double* __restrict__ c, double x, int nu)
{
    int l = bspl_interval(t, n, 4, x);
    double h[5];
    bspl_basis(t, l, x, nu ? 3 : 4, h);
    double v = 0.0;
    if (nu == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) v += c[l - 4 + i]*h[i];
        return v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int j = l - 3 + i;
        double d = 4.0*(c[j] - c[j - 1])/(t[j + 4] - t[j]);
        v += d*h[i];
    }
    return v;
}

// tensor-product spline value or partial derivative; arguments clamped to the knot domain like fpbisp
static __device__ __noinline__ double spline2d(const double* __restrict__ tx, int nx, const double* __restrict__ ty, int ny,
                                  const double* __restrict__ c, double x, double y, int dx, int dy)
{
    x = fmin(fmax(x, tx[4]), tx[nx - 5]);
    y = fmin(fmax(y, ty[4]), ty[ny - 5]);
    int lx = bspl_interval(tx, nx, 4, x), ly = bspl_interval(ty, ny, 4, y);
    int ncy = ny - 5;
    double hx[5], hy[5];
    bspl_basis(tx, lx, x, dx ? 3 : 4, hx);
    bspl_basis(ty, ly, y, dy ? 3 : 4, hy);
    double v = 0.0;
    int nbx = dx ? 4 : 5, nby = dy ? 4 : 5;
#pragma unroll 1
    for (int i = 0; i < nbx; ++i) {
        int jx = dx ? (lx - 3 + i) : (lx - 4 + i);
        double row = 0.0;
        const double hxi = (i == 0) ? hx[0] : (i == 1) ? hx[1] : (i == 2) ? hx[2] : (i == 3) ? hx[3] : hx[4];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (j < nby) {
                int jy = dy ? (ly - 3 + j) : (ly - 4 + j);
                double cc;
                if (dx) cc = 4.0*(c[jx*ncy + jy] - c[(jx - 1)*ncy + jy])/(tx[jx + 4] - tx[jx]);
                else if (dy) cc = 4.0*(c[jx*ncy + jy] - c[jx*ncy + jy - 1])/(ty[jy + 4] - ty[jy]);
                else cc = c[jx*ncy + jy];
                row += cc*hy[j];
            }
        }
        v += row*hxi;
    }
    return v;
}

// np.polyval (Horner), coefficients highest order first
__device__ __forceinline__ double polyval(const double* __restrict__ c, int n, double x)
{
    double y = 0.0;
    for (int i = 0; i < n; ++i) y = y*x + c[i];
    return y;
}

// ------------------------------------------------------------------------------------------------
// mask: Surface.mask (surface.py:235-245), RingSurface.mask (ring_surface.py:123-133),
// RectangularSurface.mask (rectangular_surface.py:100-112), SlitSurface.mask (slit_surface.py:89-102),
// FunctionSurface2D.mask (function_surface_2d.py:158-191).  Absolute coordinates.
// ------------------------------------------------------------------------------------------------
// KIND >= 0: the surface kind is a compile-time constant (the kind-specialised copies of the numeric lens step, see
// trace_step): every kind switch folds and the loop-invariant surface parameters are hoisted out of the Illinois loop.
template <int KIND = -1>
__device__ inline bool surf_mask(const KSurface& S, double x, double y)
{
    const double x0 = S.pos[0], y0 = S.pos[1];
    const int kind = (KIND >= 0) ? KIND : S.kind;
    switch (kind) {
    case OTB_SURF_RING: {
        double dx = x - x0, dy = y - y0;
        double r2 = dx*dx + dy*dy;
        double a = S.par[OTB_P_RI] - OTB_N_EPS, b = S.r + OTB_N_EPS;
        return (a*a <= r2) && (r2 <= b*b);
    }
    case OTB_SURF_RECT:
    case OTB_SURF_SLIT: {
        double xr, yr;
        rot_rc(S, OTB_P_COSM, x - x0, y - y0, xr, yr);
        double xe = S.par[OTB_P_DIMX]/2, ye = S.par[OTB_P_DIMY]/2;
        bool m = (-xe - OTB_N_EPS <= xr) && (xr <= xe + OTB_N_EPS) && (-ye - OTB_N_EPS <= yr) && (yr <= ye + OTB_N_EPS);
        if (kind == OTB_SURF_SLIT) {
            double xi = S.par[OTB_P_DIMIX]/2, yi = S.par[OTB_P_DIMIY]/2;
            bool inside = (-xi + OTB_N_EPS <= xr) && (xr <= xi - OTB_N_EPS) && (-yi + OTB_N_EPS <= yr) && (yr <= yi - OTB_N_EPS);
            m = m && !inside;
        }
        return m;
    }
    default: {
        double dx = x - x0, dy = y - y0;
        double b = S.r + OTB_N_EPS;
        bool m = dx*dx + dy*dy <= b*b;
        if (kind == OTB_SURF_FUNC && (S.flags & OTB_SF_HAS_MASK)) {
            int id = (int)S.par[OTB_P_FMASK];
            double mf;
            if (S.flags & OTB_SF_1D) {
                mf = otb_user_f1(id, sqrt(dx*dx + dy*dy));
            } else {
                double xr, yr;
                rot_rc(S, OTB_P_COSM, dx, dy, xr, yr);
                mf = otb_user_f2(id, xr, S.par[OTB_P_SIGN]*yr);
            }
            m = m && (mf != 0.0);
        }
        return m;
    }
    }
}

// ------------------------------------------------------------------------------------------------
// relative height without masking: Surface._values and overrides
// (conic_surface.py:57-68, tilted_surface.py:61-74, aspheric_surface.py:51-66,
//  function_surface_2d.py:133-156, data_surface_2d.py:130-153)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double conic_values_rel(const KSurface& S, double x, double y)
{
    double r2 = x*x + y*y;
    return S.par[OTB_P_RHO]*r2/(1 + sqrt(1 - S.par[OTB_P_KP1RHO2]*r2));
}

template <int KIND = -1>
__device__ __forceinline__ double surf_values_rel_body(const KSurface& S, const double* __restrict__ aux, double x, double y)
{
    switch ((KIND >= 0) ? KIND : S.kind) {
    case OTB_SURF_CONIC:
        return conic_values_rel(S, x, y);
    case OTB_SURF_TILTED:
        return x*S.par[OTB_P_MX] + y*S.par[OTB_P_MY];
    case OTB_SURF_ASPHERE: {
        double r = sqrt(x*x + y*y);
        double r2 = r*r;
        double z = S.par[OTB_P_RHO]*r2/(1 + sqrt(1 - S.par[OTB_P_KP1RHO2]*r2));
        z = z + polyval(aux + S.aux_off, S.aux_n0, r);
        return z - S.par[OTB_P_R];
    }
    case OTB_SURF_FUNC: {
        double sign = S.par[OTB_P_SIGN], v;
        if (S.flags & OTB_SF_1D) {
            v = otb_user_f1(S.func_id, sqrt(x*x + y*y));
        } else {
            double xr, yr;
            rot_rc(S, OTB_P_COSM, x, y, xr, yr);
            v = otb_user_f2(S.func_id, xr, sign*yr);
        }
        return sign*(v - S.par[OTB_P_OFFSET]);
    }
    case OTB_SURF_DATA: {
        double sign = S.par[OTB_P_SIGN], v;
        const double* t = aux + S.aux_off;
        if (S.flags & OTB_SF_1D) {
            v = spline1d(t, S.aux_n0, t + S.aux_n0, hypot(x, sign*y), 0);
        } else {
            double xr, yr;
            rot_rc(S, OTB_P_COSM, x, y, xr, yr);
            v = spline2d(t, S.aux_n0, t + S.aux_n0, S.aux_n1, t + S.aux_n0 + S.aux_n1, xr, sign*yr, 0, 0);
        }
        return sign*(v - S.par[OTB_P_OFFSET]);
    }
    default:
        return 0.0;
    }
}

// the one out-of-line copy of the height switch (all kinds, user functions, splines)
static __device__ __noinline__ double surf_values_rel(const KSurface* S, const double* aux, double x, double y)
{
    return surf_values_rel_body(*S, aux, x, y);
}
__device__ __forceinline__ double surf_values_rel(const KSurface& S, const double* __restrict__ aux, double x, double y)
{
    return surf_values_rel(&S, aux, x, y);
}

// Surface.values (surface.py:137-164): absolute height with the radially continued edge
template <int CAPS, int KIND = -1>
__device__ __forceinline__ double surf_values_body(const KSurface& S, const double* __restrict__ aux, double x, double y)
{
    if (S.flags & OTB_SF_FLAT) return S.z_max;
    if (CAPS == OTB_CAPS_LENS && S.kind != OTB_SURF_CONIC) return S.z_max;
    if (CAPS == OTB_CAPS_DET && S.kind != OTB_SURF_CONIC && S.kind != OTB_SURF_TILTED) return S.z_max;
    double xe = x - S.pos[0], ye = y - S.pos[1];
    if (!surf_mask<KIND>(S, x, y)) {
        if (S.flags & OTB_SF_ROTSYM) return S.pos[2] + S.par[OTB_P_EDGEZ];
        double r = S.r - OTB_N_EPS;
        double phi = atan2(ye, xe);
        xe = r*cos(phi);
        ye = r*sin(phi);
    }
    // ONE inlined copy of the height expression(s) in here
    if (CAPS == OTB_CAPS_LENS) return S.pos[2] + conic_values_rel(S, xe, ye);
    if (CAPS == OTB_CAPS_DET)
        return S.pos[2] + ((S.kind == OTB_SURF_TILTED) ? xe*S.par[OTB_P_MX] + ye*S.par[OTB_P_MY] : conic_values_rel(S, xe, ye));
    return S.pos[2] + surf_values_rel_body<KIND>(S, aux, xe, ye);
}
static __device__ __noinline__ double surf_values_full(const KSurface* S, const double* aux, double x, double y)
{
    return surf_values_body<OTB_CAPS_FULL>(*S, aux, x, y);
}
template <int CAPS>
__device__ __forceinline__ double surf_values(const KSurface& S, const double* __restrict__ aux, double x, double y)
{
    if (CAPS == OTB_CAPS_FULL) return surf_values_full(&S, aux, x, y);
    return surf_values_body<CAPS>(S, aux, x, y);
}

// ------------------------------------------------------------------------------------------------
// normals: Surface.normals (surface.py:247-285), ConicSurface.normals (conic_surface.py:70-124),
// TiltedSurface.normals (tilted_surface.py:76-89), FunctionSurface2D.normals
// (function_surface_2d.py:193-253), DataSurface2D.normals (data_surface_2d.py:155-196)
// ------------------------------------------------------------------------------------------------
// HOT = 1: the inlined instantiation of the numeric-surface lens step in the trace loop (asphere / function / data
// surfaces only, see trace_step): no conic code, heights evaluated inline at ONE code site per loop.
template <int CAPS, int HOT = 0, int KIND = -1>
__device__ inline V3 surf_normal(const KSurface& S, const double* __restrict__ aux, double x, double y)
{
    const int k = (KIND >= 0) ? KIND : S.kind;
    if ((S.flags & OTB_SF_FLAT) && k != OTB_SURF_TILTED) return v3(0.0, 0.0, 1.0);
    if (!surf_mask<KIND>(S, x, y)) return v3(0.0, 0.0, 1.0);
    if (CAPS == OTB_CAPS_LENS && k != OTB_SURF_CONIC) return v3(0.0, 0.0, 1.0);
    const double x0 = S.pos[0], y0 = S.pos[1];
    const double dx = x - x0, dy = y - y0;

    if (!HOT && k == OTB_SURF_CONIC) {
        double rho = S.par[OTB_P_RHO];
        if (S.par[OTB_P_K] == 0.0) {
            double rho2 = S.par[OTB_P_RHO2];
            return v3(-rho*dx, -rho*dy, sqrt(1 - rho2*(dx*dx) - rho2*(dy*dy)));
        }
        // cos(phi), sin(phi) of phi = atan2(dy, dx) (conic_surface.py:104-110) taken as dx/r, dy/r: the same
        // numbers to an ulp without three transcendental calls per ray; at the vertex phi = atan2(0, 0) = 0
        double r = sqrt(dx*dx + dy*dy);
        double n_r = -rho*r/sqrt(1 - S.par[OTB_P_KRHO2]*(r*r));
        double c = 1.0, sn = 0.0;
        if (r != 0.0) {
            const double yr = rcp_seq(r);
            c = div_seq(dx, r, yr);
            sn = div_seq(dy, r, yr);
        }
        return v3(n_r*c, n_r*sn, sqrt(1 - n_r*n_r));
    }
    if (!HOT && k == OTB_SURF_TILTED) return v3(S.par[OTB_P_NX], S.par[OTB_P_NY], S.par[OTB_P_NZ]);

    const bool analytic = (k == OTB_SURF_ASPHERE) || (k == OTB_SURF_DATA) || (k == OTB_SURF_FUNC && (S.flags & OTB_SF_HAS_DERIV));
    if (analytic) {
        double nxn, nyn;
        if (S.flags & OTB_SF_1D) {
            double phi = atan2(dy, dx), nr;
            if (k == OTB_SURF_ASPHERE) {
                double rm = sqrt(dx*dx + dy*dy);
                double fr = rm*S.par[OTB_P_RHO]/sqrt(1 - S.par[OTB_P_KP1RHO2]*(rm*rm));
                nr = fr + polyval(aux + S.aux_off + S.aux_n0, S.aux_n1, rm);
            } else if (k == OTB_SURF_FUNC) {
                nr = S.par[OTB_P_SIGN]*otb_user_f1((int)S.par[OTB_P_FDERIV], sqrt(dx*dx + dy*dy));
            } else {
                const double* t = aux + S.aux_off;
                nr = S.par[OTB_P_SIGN]*spline1d(t, S.aux_n0, t + S.aux_n0, hypot(dx, dy), 1);
            }
            nxn = nr*cos(phi);
            nyn = nr*sin(phi);
        } else {
            double sign = S.par[OTB_P_SIGN], a, b;
            if (k == OTB_SURF_FUNC) {
                otb_user_d2((int)S.par[OTB_P_FDERIV], dx, sign*dy, &a, &b);   // unrotated, function_surface_2d.py:234
                a = a*sign;
            } else {
                double xr, yr;
                rot_rc(S, OTB_P_COSM, dx, dy, xr, yr);
                const double* t = aux + S.aux_off;
                const double *ty = t + S.aux_n0, *c = t + S.aux_n0 + S.aux_n1;
                a = spline2d(t, S.aux_n0, ty, S.aux_n1, c, xr, sign*yr, 1, 0)*sign;
                b = spline2d(t, S.aux_n0, ty, S.aux_n1, c, xr, sign*yr, 0, 1);
            }
            rot_rc(S, OTB_P_COSP, a, b, nxn, nyn);
        }
        return unit3(v3(-nxn, -nyn, 1.0));
    }

    // central differences (surface.py:266-283)
    double eps = S.par[OTB_P_FDEPS];
    V3 n;
    if (HOT) {
        // the four heights through one inlined copy of the height switch (rolled loop)
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3_ = 0.0;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const double xx = (q == 0) ? dx - eps : (q == 1) ? dx + eps : dx;
            const double yy = (q == 2) ? dy - eps : (q == 3) ? dy + eps : dy;
            const double v = surf_values_rel_body<KIND>(S, aux, xx, yy);
            v0 = (q == 0) ? v : v0;
            v1 = (q == 1) ? v : v1;
            v2 = (q == 2) ? v : v2;
            v3_ = (q == 3) ? v : v3_;
        }
        n.x = v0 - v1;
        n.y = v2 - v3_;
    } else {
        n.x = surf_values_rel(S, aux, dx - eps, dy) - surf_values_rel(S, aux, dx + eps, dy);
        n.y = surf_values_rel(S, aux, dx, dy - eps) - surf_values_rel(S, aux, dx, dy + eps);
    }
    n.z = 2*eps;
    return unit3(n);
}

// ------------------------------------------------------------------------------------------------
// intersection
// ------------------------------------------------------------------------------------------------
// Surface._find_hit_handle_abnormal (surface.py:436-479).  dz = h.p.z - values(h.p.x, h.p.y), supplied by the caller.
__device__ __forceinline__ void handle_abnormal_dz(const KSurface& S, const V3& p, const V3& s, HitResult& h, double dz)
{
    bool dev = fabs(dz) > OTB_C_EPS;
    bool beh = p.z > S.z_max + OTB_N_EPS;
    bool neg = h.p.z < p.z - OTB_C_EPS;
    bool bet = (neg || dev) && !beh;
    if (bet) {
        double tnm = (S.z_max - p.z)/s.z;
        h.p = along(p, s, tnm);
        h.hit = false;
    }
    if (beh) {
        h.p = p;
        h.hit = false;
    }
}

template <int CAPS>
__device__ __forceinline__ void handle_abnormal(const KSurface& S, const double* __restrict__ aux, const V3& p, const V3& s, HitResult& h)
{
    double zs = surf_values<CAPS>(S, aux, h.p.x, h.p.y);
    handle_abnormal_dz(S, p, s, h, h.p.z - zs);
}

// Surface.find_hit (surface.py:307-414): plane for flat surfaces, Illinois regula falsi otherwise.
// `status` receives OTB_STATUS_TIMEOUT when the 200-iteration limit is reached (surface.py:403).
// The surface height is evaluated at ONE code site: the two bracket ends (surface.py:340-347), every secant point
// (:365-367) and the final deviation check of _find_hit_handle_abnormal (:452, whose argument is the last point
// evaluated, so its height difference is the value already at hand) — same operations on the same operands as the
// reference, a third of the code.
template <int CAPS, int HOT = 0, int KIND = -1>
__device__ inline HitResult find_hit_numeric(const KSurface& S, const double* __restrict__ aux, const V3& p, const V3& s, int* status)
{
    HitResult h;
    h.ill = false;
    if (!HOT && (S.flags & OTB_SF_FLAT)) {
        double t = (S.pos[2] - p.z)/s.z;
        h.p = along(p, s, t);
        h.hit = surf_mask(S, h.p.x, h.p.y);
        handle_abnormal<CAPS>(S, aux, p, s, h);
        return h;
    }
    if (CAPS == OTB_CAPS_LENS) {      // not reachable for scenes admitted to this instantiation
        h.p = p;
        h.hit = false;
        return h;
    }
    double t1 = (S.z_min - OTB_C_EPS/10 - p.z)/s.z;
    double t2 = (S.z_max + OTB_C_EPS/10 - p.z)/s.z;
    if (t1 < 0) t1 = -OTB_C_EPS;
    bool w = true;
    if (!finite_d(t1) || !finite_d(t2)) w = false;
    if ((t2 - t1) < OTB_C_EPS) w = false;
    double f1 = 0.0, f2 = 0.0, dz = 0.0;
    int phase = 0;                    // 0: bracket start, 1: bracket end, 2: secant points
    int it = 1;
    h.p = v3(0.0, 0.0, 0.0);
#pragma unroll 1
    for (;;) {
        const double ts = (phase == 0) ? t1 : (phase == 1) ? t2 : t1 - f1/(f2 - f1)*(t2 - t1);
        const V3 pl = along(p, s, ts);
        const double fts = pl.z - (HOT ? surf_values_body<CAPS, KIND>(S, aux, pl.x, pl.y) : surf_values<CAPS>(S, aux, pl.x, pl.y));
        if (phase == 0) {
            f1 = fts;
            h.p = pl;                 // kept when the bracket is degenerate (surface.py:352)
            dz = fts;
            phase = 1;
            continue;
        }
        if (phase == 1) {
            f2 = fts;
            h.ill = f1*f2 > 0;
            phase = 2;
            if (!w) break;
            continue;
        }
        const double prod = fts*f2;
        if (prod < 0) {            // case 1: [t2, ts]
            t1 = t2; t2 = ts; f1 = f2; f2 = fts;
        } else if (prod > 0) {     // case 2: [t1, ts], Illinois factor 0.5 on the retained end
            t2 = ts; f1 = 0.5*f1; f2 = fts;
        } else if (prod == 0) {    // case 3: exact root
            t1 = ts; t2 = ts; f1 = fts; f2 = fts;
        }
        if (fabs(t2 - t1) < OTB_C_EPS/10) {
            h.p = pl;
            dz = fts;
            break;
        }
        if (it == 200) {
            atomicOr(status, OTB_STATUS_TIMEOUT);
            h.p = pl;
            dz = fts;
            break;
        }
        ++it;
    }
    h.hit = surf_mask<KIND>(S, h.p.x, h.p.y);
    handle_abnormal_dz(S, p, s, h, dz);
    return h;
}

// ConicSurface.find_hit (conic_surface.py:126-203)
__device__ inline HitResult find_hit_conic(const KSurface& S, const V3& p, const V3& s)
{
    HitResult h;
    h.ill = false;
    const double ox = p.x - S.pos[0], oy = p.y - S.pos[1], oz = p.z - S.pos[2];
    const double k = S.par[OTB_P_K], kp1 = S.par[OTB_P_KP1];
    const double A = (k != 0.0) ? 1 + k*(s.z*s.z) : 1.0;
    const double B = s.x*ox + s.y*oy + s.z*(oz*kp1 - S.par[OTB_P_INVRHO]);
    const double Cc = oy*oy + ox*ox + oz*(oz*kp1 - S.par[OTB_P_TWOINVRHO]);
    const double D = sqrt(B*B - Cc*A);
    // sphere: A is the literal 1.0 and x/1.0 == x exactly, so the two divisions are skipped
    const bool unitA = (A == 1.0);
    const double t1 = unitA ? (-B - D) : (-B - D)/A, t2 = unitA ? (-B + D) : (-B + D)/A;
    const double z = p.z;
    const double z1 = z + s.z*t1, z2 = z + s.z*t2;
    const double z_min = S.z_min - OTB_N_EPS, z_max = S.z_max + OTB_N_EPS;
    const bool c1 = (z_min <= z1) && (z1 <= z_max) && (z1 >= z);
    const bool c2 = (z_min <= z2) && (z2 <= z_max) && (z2 >= z) && (t2 < t1);
    double t = (c1 && !c2) ? t1 : t2;
    h.p = along(p, s, t);
    h.hit = surf_mask(S, h.p.x, h.p.y);
    if (A == 0.0 && B != 0.0) {
        double tl = -Cc/(2*B);
        h.p = along(p, s, tl);
        h.hit = surf_mask(S, h.p.x, h.p.y);
    }
    bool nh = !h.hit || !finite_d(D) || (A == 0.0 && B == 0.0) || (h.p.z < z_min) || (h.p.z > z_max);
    if (nh) {
        double tnh = (S.z_max - p.z)/s.z;
        h.p = along(p, s, tnh);
        h.hit = false;
    }
    if (z > S.z_max) {
        h.p = p;
        h.hit = false;
    }
    return h;
}

// TiltedSurface.find_hit (tilted_surface.py:91-123): analytic plane hit; rays that miss the disc go through the
// numeric finder (radially continued edge).  Shares the ONE inlined copy of find_hit_numeric with the other kinds.
template <int CAPS, int HOT = 0, int KIND = -1>
__device__ inline HitResult surf_find_hit(const KSurface& S, const double* __restrict__ aux, const V3& p, const V3& s, int* status)
{
    if (HOT) return find_hit_numeric<CAPS, 1, KIND>(S, aux, p, s, status);
    if (S.kind == OTB_SURF_CONIC) return find_hit_conic(S, p, s);
    const bool tilted = (CAPS != OTB_CAPS_LENS) && S.kind == OTB_SURF_TILTED;
    HitResult h;
    h.hit = false;
    if (tilted) {
        const V3 n = v3(S.par[OTB_P_NX], S.par[OTB_P_NY], S.par[OTB_P_NZ]);
        double t_denom = dot3(s, n);
        bool nz = t_denom != 0;
        V3 d = v3(S.pos[0] - p.x, S.pos[1] - p.y, S.pos[2] - p.z);
        double t = dot3(d, n)/(nz ? t_denom : 1e-12);
        h.p = along(p, s, t);
        h.hit = surf_mask(S, h.p.x, h.p.y) && nz;
        h.ill = false;
    }
    if (!h.hit) h = find_hit_numeric<CAPS>(S, aux, p, s, status);
    if (tilted) handle_abnormal<CAPS>(S, aux, p, s, h);
    return h;
}

// SphericalSurface.sphere_projection (spherical_surface.py:36-97), in place on (x, y) given z
__device__ __forceinline__ void sphere_project(const KSurface& S, int method, double& x, double& y, double z)
{
    if (method == OTB_PROJ_NONE || method == OTB_PROJ_ORTHOGRAPHIC) return;
    const double R = S.par[OTB_P_R];
    const double dx = x - S.pos[0], dy = y - S.pos[1];
    const double zm = S.pos[2] + R;
    const double sgn = (R > 0) ? 1.0 : ((R < 0) ? -1.0 : 0.0);
    if (method == OTB_PROJ_EQUIDISTANT) {
        double r = sqrt(dx*dx + dy*dy);
        double theta = -sgn*atan(r/(z - zm));
        double phi = atan2(dy, dx);
        x = theta*cos(phi);
        y = theta*sin(phi);
    } else if (method == OTB_PROJ_STEREOGRAPHIC) {
        double r = sqrt(dx*dx + dy*dy);
        double theta = 1.5707963267948966 - atan(r/(z - zm));
        double phi = atan2(dy, dx);
        double rr = -2*sgn*tan(0.7853981633974483 - theta/2);
        x = rr*cos(phi);
        y = rr*sin(phi);
    } else {   // Equal-Area
        double x_ = dx/fabs(R), y_ = dy/fabs(R), z_ = (z - zm)/R;
        double f = sqrt(2/(1 - z_));
        x = f*x_;
        y = f*y_;
    }
}

