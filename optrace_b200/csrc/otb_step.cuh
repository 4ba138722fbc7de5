// otb_step.cuh — one ray crossing one tracing surface: the body of the reference's sub_trace loop
// (raytracer.py:307-397) restated per ray, with all ray state in registers.
#pragma once
#include "otb_common.cuh"
#include "otb_surfaces.cuh"
#include "otb_media.cuh"
#include "otb_rng.cuh"

struct RayState {
    V3 p;          // position at section i
    V3 s;          // current unit direction
    float pol[3];  // polarisation at section i (float32 like RayStorage.pol_list)
    float w;       // weight at section i (float32 like RayStorage.w_list)
    float wl;      // wavelength in nm (float32 like RayStorage.wl_list)
    double n;      // refraction index of the current medium = n_list[:, i]
};

// info-message predicates of one step (Raytracer.INFOS, raytracer.py:43-48); booked by the caller
struct StepFlags {
    bool ill, absorb_missing, tir, outline, hurb_neg;
};

#define OTB_INV_SQRT2 (1.0/1.4142135623730951)   // 1/np.sqrt(2)

#include "otb_fast.cuh"

#ifndef OTB_STEP_OOL
#define OTB_STEP_OOL 1
#endif

// Raytracer.__compute_polarization (raytracer.py:831-879): returns amplitude components and writes the
// new polarisation when the direction changed.
template <bool POL>
__device__ __forceinline__ void compute_polarization(const V3& s, const V3& s_, const float* pol_i, float* pol_n,
                                                     double& A_ts, double& A_tp)
{
    if (!POL) {
        A_ts = OTB_INV_SQRT2;
        A_tp = OTB_INV_SQRT2;
        return;
    }
    const bool changed = (s.x != s_.x) || (s.y != s_.y) || (s.z != s_.z);
    if (!changed) {
        A_ts = OTB_INV_SQRT2;
        A_tp = OTB_INV_SQRT2;
        return;
    }
    const V3 ps = unit3(cross3(s_, s));
    const V3 pp = cross3(ps, s);
    const V3 pol = v3((double)pol_i[0], (double)pol_i[1], (double)pol_i[2]);
    A_ts = dot3(ps, pol);
    A_tp = dot3(pp, pol);
    const V3 pp_ = cross3(ps, s_);
    pol_n[0] = (float)(ps.x*A_ts + pp_.x*A_tp);
    pol_n[1] = (float)(ps.y*A_ts + pp_.y*A_tp);
    pol_n[2] = (float)(ps.z*A_ts + pp_.z*A_tp);
}

// Raytracer.__outline_intersection (raytracer.py:666-718) for one ray: returns true when clipped
__device__ __forceinline__ bool outline_clip(const double* __restrict__ o, const V3& p_i, const V3& s, V3& p_n)
{
    const bool inside = (o[0] < p_n.x) && (p_n.x < o[1]) && (o[2] < p_n.y) && (p_n.y < o[3]) && (o[4] < p_n.z) && (p_n.z < o[5]);
    if (inside) return false;
    double t = nan("");
    const double P[3] = {p_i.x, p_i.y, p_i.z}, S[3] = {s.x, s.y, s.z};
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double T = (o[j] - P[j >> 1])/S[j >> 1];
        if (T > 0) t = (t != t) ? T : fmin(t, T);      // smallest positive parameter, NaN/<=0 ignored (np.nanmin)
    }
    p_n = along(p_i, s, t);
    return true;
}

// HURB geometry: RingSurface.hurb_props (ring_surface.py:88-121), SlitSurface.hurb_props (slit_surface.py:65-87)
__device__ __forceinline__ void hurb_props(const KSurface& S, double x, double y, double& a_, double& b_, V3& b, bool& inside)
{
    const double dx = x - S.pos[0], dy = y - S.pos[1];
    if (S.kind == OTB_SURF_RING) {
        double r = sqrt(dx*dx + dy*dy);
        double theta = atan2(dy, dx);
        double R = S.par[OTB_P_RI];
        inside = r < R;
        b_ = R - r;
        a_ = sqrt(b_*R);
        b = v3(cos(theta), sin(theta), 0.0);
    } else {
        double xr, yr;
        rot_rc(S, OTB_P_COSM, dx, dy, xr, yr);
        a_ = S.par[OTB_P_DIMIY]/2 - fabs(yr);
        b_ = S.par[OTB_P_DIMIX]/2 - fabs(xr);
        inside = (a_ > 0) && (b_ > 0);
        b = v3(S.par[OTB_P_COSP], S.par[OTB_P_SINP], 0.0);
    }
}

// One sequential step.  On entry `r` holds section i, on exit section i+1 (p, w, pol, n) and the new direction.
// za, zb: standard normal deviates for HURB (only read when the step bends rays).
// HOT = 1: lens front / back at a non-flat asphere / function / data surface only (the inlined numeric-surface step of
// trace_step); everything else of the step is compiled out.
template <bool POL, int CAPS, int HOT = 0, int KIND = -1>
__device__ __forceinline__ void trace_step_full(const KScene& sc, const double* __restrict__ aux, const OtbStep& st, RayState& r,
                                                StepFlags& fl, double za, double zb, int* status)
{
    const KSurface& S = sc.surf[st.surface];
    fl.ill = fl.absorb_missing = fl.tir = fl.outline = fl.hurb_neg = false;

    const bool hw = r.w > 0.0f;
    const V3 p_i = r.p;
    V3 p_n = r.p;
    float w_n = r.w;
    float pol_n[3] = {r.pol[0], r.pol[1], r.pol[2]};
    bool hit = false;

    if (hw) {
        HitResult h = surf_find_hit<CAPS, HOT, KIND>(S, aux, r.p, r.s, status);
        p_n = h.p;
        hit = h.hit;
        fl.ill = h.ill;
    }
    const bool hwh = hw && hit, hwnh = hw && !hit;

    if (HOT || st.role <= OTB_STEP_IDEAL_LENS) {
        // ---- Lens front / back / ideal lens (raytracer.py:314-370) ----
        if (hwnh) {
            w_n = 0.0f;
            if (st.role == OTB_STEP_LENS_BACK) p_n = p_i;     // absorbed at the lens front (raytracer.py:354)
            fl.absorb_missing = true;
        }
        const double n2 = medium_n(sc.media[st.medium_after], aux, r.wl);
        if (n2 < 1.0) atomicOr(status, OTB_STATUS_NBELOW1);
        if (hwh) {
            if (!HOT && st.role == OTB_STEP_IDEAL_LENS) {
                // Raytracer.__refraction_ideal_lens (raytracer.py:720-759)
                const V3 s0 = r.s;
                const double f = 1000/st.D;
                const double fsz = f/s0.z;
                V3 sn = v3(s0.x*fsz - (p_n.x - S.pos[0]), s0.y*fsz - (p_n.y - S.pos[1]), f);
                sn = unit3(sn);
                const double sg = (f > 0) ? 1.0 : ((f < 0) ? -1.0 : 0.0);
                r.s = v3(sn.x*sg, sn.y*sg, sn.z*sg);
                double a, b;
                compute_polarization<POL>(s0, r.s, r.pol, pol_n, a, b);
            } else {
                // Raytracer.__refraction (raytracer.py:761-829)
                const V3 nrm = surf_normal<CAPS, HOT, KIND>(S, aux, p_n.x, p_n.y);
                const double n1 = r.n;
                const double ns = dot3(nrm, r.s);
                const double N = n1/n2;
                const double W = sqrt(1 - (N*N)*(1 - ns*ns));
                const double q = N*ns - W;
                const V3 s_ = v3(r.s.x*N - nrm.x*q, r.s.y*N - nrm.y*q, r.s.z*N - nrm.z*q);
                double A_ts, A_tp;
                compute_polarization<POL>(r.s, s_, r.pol, pol_n, A_ts, A_tp);
                const double n1ca = n1*ns, n2cb = n2*W;
                const double ts = 2*n1ca/(n1ca + n2cb);
                const double tp = 2*n1ca/(n2*ns + n1*W);
                const double ats = A_ts*ts, atp = A_tp*tp;
                double T = n2cb/n1ca*(ats*ats + atp*atp);
                if (!finite_d(W)) {
                    T = 0.0;
                    fl.tir = true;
                }
                w_n = (float)((double)r.w*T);
                r.s = s_;
            }
        }
        if (hwnh) {
            if (outline_clip(sc.outline, p_i, r.s, p_n)) {
                w_n = 0.0f;
                fl.outline = true;
            }
        }
        r.n = n2;
    } else {
        // ---- Filter / Aperture (raytracer.py:372-391) ----
        if (st.role == OTB_STEP_FILTER) {
            if (hwh) w_n = filter_apply(sc.filters[st.filter], aux, r.wl, r.w);
        } else {
            if (hwh) w_n = 0.0f;
            if (CAPS == OTB_CAPS_FULL && st.hurb) {
                // Raytracer.__hurb (raytracer.py:417-490)
                const V3 s0 = r.s;
                if (hwnh) {
                    double a_, b_;
                    V3 b;
                    bool inside;
                    hurb_props(S, p_n.x, p_n.y, a_, b_, b, inside);
                    if (inside) {
                        const V3 a = v3(-b.y, b.x, 0.0);
                        const double da = dot3(r.s, a), db = dot3(r.s, b);
                        const double cos_psi_a = sqrt(1 - da*da), cos_psi_b = sqrt(1 - db*db);
                        const double k = 6.283185307179586*r.n/(double)__fmul_rn(r.wl, 1e-9f);
                        const double tan_sig_b = sc.hurb_factor/(2*b_*cos_psi_b*1e-3*k);
                        const double tan_sig_a = sc.hurb_factor/(2*a_*cos_psi_a*1e-3*k);
                        const double tan_tha = fabs(tan_sig_a)*za, tan_thb = fabs(tan_sig_b)*zb;
                        const V3 sa = unit3(cross3(b, r.s));
                        const V3 sb = cross3(r.s, sa);
                        const V3 sab = v3(r.s.x + sa.x*tan_tha + sb.x*tan_thb, r.s.y + sa.y*tan_tha + sb.y*tan_thb,
                                          r.s.z + sa.z*tan_tha + sb.z*tan_thb);
                        r.s = unit3(sab);
                        double aa, bb;
                        compute_polarization<POL>(s0, r.s, r.pol, pol_n, aa, bb);
                    }
                }
                if (r.s.z < 0) {             // all rays, alive or not (raytracer.py:484-486)
                    w_n = 0.0f;
                    fl.hurb_neg = true;
                }
            }
        }
        if (hwnh) {
            if (outline_clip(sc.outline, p_i, r.s, p_n)) {
                w_n = 0.0f;
                fl.outline = true;
            }
        }
    }
    r.p = p_n;
    r.w = w_n;
    if (POL) {
        r.pol[0] = pol_n[0];
        r.pol[1] = pol_n[1];
        r.pol[2] = pol_n[2];
    }
}

// out-of-line copy of the full step for the rays the branch-free main path hands back (one copy per kernel, not
// one per unrolled step).  State travels by value so that the caller's copy stays in registers on the main path.
struct StepIO {
    RayState r;
    StepFlags fl;
};

template <bool POL, int CAPS>
__device__ __noinline__ StepIO trace_step_slow(const KScene* sc, const double* aux, const OtbStep* st, RayState r,
                                               double za, double zb, int* status)
{
    StepIO o;
    o.r = r;
    trace_step_full<POL, CAPS>(*sc, aux, *st, o.r, o.fl, za, zb, status);
    return o;
}

// One sequential step: spherical lens surfaces take the branch-free main path (otb_fast.cuh) and fall back to
// the full step per ray; everything else runs the full step.  With OTB_STEP_OOL the full step exists only as
// the out-of-line copy: the step loop then holds the straight-line path and ONE call site, which keeps the
// loop-carried ray state in fixed registers (no copies where the paths merge) and the hot loop small.
// func_spec (store-mode trace kernel only): function surfaces take the kind-specialised copy of the numeric step
template <bool POL, int CAPS>
__device__ __forceinline__ void trace_step(const KScene& sc, const double* __restrict__ aux, const OtbStep& st, RayState& r,
                                           StepFlags& fl, double za, double zb, int* status, const bool func_spec = false)
{
    // A warp whose 32 rays are all absorbed (w == 0) has nothing to intersect or refract: dead rays repeat their
    // position, polarisation and zero weight in every later section (raytracer.py:309-312) and only follow the media
    // (n_list is written for every ray, raytracer.py:343, 362).  Coherent bundles (OtbSource.coherent) make such
    // warps the rule behind a stop instead of the exception.
    if (!__any_sync(0xffffffffu, r.w > 0.0f)) {
        fl.ill = fl.absorb_missing = fl.tir = fl.outline = fl.hurb_neg = false;
        if (st.role <= OTB_STEP_IDEAL_LENS) {
            const double n2 = medium_n(sc.media[st.medium_after], aux, (double)r.wl);
            if (n2 < 1.0) atomicOr(status, OTB_STATUS_NBELOW1);
            r.n = n2;
        } else if (CAPS == OTB_CAPS_FULL && st.role == OTB_STEP_APERTURE && st.hurb) {
            fl.hurb_neg = r.s.z < 0;          // booked for every ray, alive or not (raytracer.py:484-486)
        }
        return;
    }
#ifndef OTB_NO_FAST_PATH
    const KSurface& S = sc.surf[st.surface];
    bool done = false;
    if (st.role <= OTB_STEP_LENS_BACK && S.kind == OTB_SURF_CONIC) {
        if (sc.arithmetic == OTB_ARITH_RELAXED) {
            if (S.par[OTB_P_K] == 0.0) done = relaxed_conic_lens_step<POL, true>(sc, aux, st, S, r, fl, status);
            else done = relaxed_conic_lens_step<POL, false>(sc, aux, st, S, r, fl, status);
        } else {
            if (S.par[OTB_P_K] == 0.0) done = fast_conic_lens_step<POL, true>(sc, aux, st, S, r, fl, status);
            else done = fast_conic_lens_step<POL, false>(sc, aux, st, S, r, fl, status);
        }
    } else if (st.role == OTB_STEP_APERTURE && !st.hurb && (S.flags & OTB_SF_FLAT) && !(S.flags & OTB_SF_ROTATED)
               && (S.kind == OTB_SURF_CIRCLE || S.kind == OTB_SURF_RECT || S.kind == OTB_SURF_RING)) {
        done = fast_flat_aperture_step(sc, S, r, fl);
    } else if (CAPS == OTB_CAPS_FULL && st.role <= OTB_STEP_LENS_BACK && !(S.flags & OTB_SF_FLAT)
               && (S.kind == OTB_SURF_FUNC || S.kind == OTB_SURF_DATA || S.kind == OTB_SURF_ASPHERE)) {
        // numeric surfaces: for them the full step IS the hot path — inlined here (scene in the constant bank, no
        // call per height evaluation) with everything but the lens branch compiled out
        // ... and one copy per surface kind: the kind switches fold and the surface parameters stay in registers
        // across the Illinois iterations instead of being re-read from the constant bank
        // The specialised copy exists for function surfaces and is taken only in scenes whose numeric surfaces are ALL
        // function surfaces (func_spec, decided once per kernel): cosine_surfaces 7.6 -> 6.7 ms; in a scene that mixes
        // kinds every additional copy in use costs instruction-cache hits (zoo_numeric 89 -> 125 ms when used there).
        if (func_spec && S.kind == OTB_SURF_FUNC) trace_step_full<POL, CAPS, 1, OTB_SURF_FUNC>(sc, aux, st, r, fl, za, zb, status);
        else trace_step_full<POL, CAPS, 1>(sc, aux, st, r, fl, za, zb, status);
        done = true;
    }
#if OTB_STEP_OOL
    if (!done) {
        const StepIO o = trace_step_slow<POL, CAPS>(&sc, aux, &st, r, za, zb, status);
        r = o.r;
        fl = o.fl;
    }
#else
    if (done) return;
    if ((st.role <= OTB_STEP_LENS_BACK && S.kind == OTB_SURF_CONIC) || st.role == OTB_STEP_APERTURE) {
        const StepIO o = trace_step_slow<POL, CAPS>(&sc, aux, &st, r, za, zb, status);
        r = o.r;
        fl = o.fl;
        return;
    }
    trace_step_full<POL, CAPS>(sc, aux, st, r, fl, za, zb, status);
#endif
#else
    trace_step_full<POL, CAPS>(sc, aux, st, r, fl, za, zb, status);
#endif
}

// warp-aggregated message booking: one shared-memory atomic per warp and message type; the common case
// (no message in the whole warp) costs one vote
__device__ __forceinline__ void book(int* smsgs, int slot, bool pred)
{
    unsigned b = __ballot_sync(0xffffffffu, pred);
    if (b && (threadIdx.x & 31) == 0) atomicAdd(&smsgs[slot], __popc(b));
}

// section indices as in raytracer.py:318, 323, 486 (i+1) vs :718, 826 (i)
__device__ __forceinline__ void book_step(int* smsgs, int nt, int i, bool valid, const StepFlags& fl)
{
    const bool any = valid && (fl.ill || fl.absorb_missing || fl.tir || fl.outline || fl.hurb_neg);
    if (__any_sync(0xffffffffu, any)) {
        book(smsgs, OTB_MSG_ILL_COND*nt + i + 1, valid && fl.ill);
        book(smsgs, OTB_MSG_ABSORB_MISSING*nt + i + 1, valid && fl.absorb_missing);
        book(smsgs, OTB_MSG_TIR*nt + i, valid && fl.tir);
        book(smsgs, OTB_MSG_OUTLINE*nt + i, valid && fl.outline);
        book(smsgs, OTB_MSG_HURB_NEG*nt + i + 1, valid && fl.hurb_neg);
    }
}

// ---- scene specialisation (optrace_b200/specialise.py) ---------------------------------------------------
// A specialised build defines OTB_SPEC_SCENE_H: the scene becomes a __device__ const object, the step loop is
// unrolled and all kind/role/flag dispatch folds at compile time; only one kernel instantiation is generated.
#ifdef OTB_SPEC_SCENE_H
#include OTB_SPEC_SCENE_H
#define OTB_SPEC 1
static __device__ const KScene K_SPEC = OTB_SPEC_SCENE_INIT;
static const KScene K_SPEC_HOST = OTB_SPEC_SCENE_INIT;
#else
#define OTB_SPEC 0
#endif

// resident blocks per SM the register allocation aims at (128 threads per block): measured optimum per variant
#if OTB_SPEC && defined(OTB_SPEC_MINBLOCKS)
#define OTB_MINBLOCKS(CAPS) OTB_SPEC_MINBLOCKS
#else
#ifndef OTB_LENS_MINBLOCKS
#define OTB_LENS_MINBLOCKS 4
#endif
#define OTB_MINBLOCKS(CAPS) ((CAPS) == OTB_CAPS_LENS ? OTB_LENS_MINBLOCKS : 3)
#endif
