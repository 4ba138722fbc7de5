// otb_media.cuh — wavelength-dependent scalars: refraction index models, filter transmission,
// linear table interpolation (np.interp semantics) and the CIE observer lookup.
#pragma once
#include "otb_common.cuh"

// np.interp(x, xp, fp, left=0, right=0) for one sample (numpy/_core/src/multiarray/compiled_base.c,
// arr_interp): binary search for xp[j] <= x < xp[j+1], slope*(x - xp[j]) + fp[j], exact hits return fp[j].
__device__ inline double interp_lr0(const double* __restrict__ xp, const double* __restrict__ fp, int n, double x)
{
    if (!(x >= xp[0]) || !(x <= xp[n - 1])) return (x != x) ? x : 0.0;   // NaN propagates, outside -> 0
    if (x == xp[n - 1]) return fp[n - 1];
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x >= xp[mid]) lo = mid; else hi = mid;
    }
    if (xp[lo] == x) return fp[lo];
    double slope = (fp[lo + 1] - fp[lo])/(xp[lo + 1] - xp[lo]);
    double r = slope*(x - xp[lo]) + fp[lo];
    if (r != r) {
        r = slope*(x - xp[lo + 1]) + fp[lo + 1];
        if (r != r && fp[lo] == fp[lo + 1]) r = fp[lo];
    }
    return r;
}

// RefractionIndex.__call__ (refraction_index.py:62-169).  wl in nm (the ray path passes its float32 wavelength upcast exactly, :70).
__device__ inline double medium_n(const OtbMedium& M, const double* __restrict__ aux, double wl)
{
    const double* c = M.c;
    const double l = wl*1e-3;
    const double w2 = l*l;     // (wl*1e-3)**2
    switch (M.model) {
    case OTB_N_CONSTANT: return c[0];
    case OTB_N_ABBE: return c[0] + c[1]/(w2 - c[2]);
    case OTB_N_CAUCHY: return c[0] + c[1]/w2 + c[2]/(w2*w2) + c[3]/pow(w2, 3.0);
    case OTB_N_CONRADY: return c[0] + c[1]/l + c[2]/pow(l, 3.5);
    case OTB_N_SELLMEIER1:
        return sqrt(1 + c[0]*w2/(w2 - c[1]) + c[2]*w2/(w2 - c[3]) + c[4]*w2/(w2 - c[5]));
    case OTB_N_SELLMEIER2:
        return sqrt(1 + c[0] + c[1]*w2/(w2 - c[2]*c[2]) + c[3]/(w2 - c[4]*c[4]));
    case OTB_N_SELLMEIER3:
        return sqrt(1 + c[0]*w2/(w2 - c[1]) + c[2]*w2/(w2 - c[3]) + c[4]*w2/(w2 - c[5]) + c[6]*w2/(w2 - c[7]));
    case OTB_N_SELLMEIER4:
        return sqrt(c[0] + c[1]*w2/(w2 - c[2]) + c[3]*w2/(w2 - c[4]));
    case OTB_N_SELLMEIER5:
        return sqrt(1 + c[0]*w2/(w2 - c[1]) + c[2]*w2/(w2 - c[3]) + c[4]*w2/(w2 - c[5]) + c[6]*w2/(w2 - c[7])
                    + c[8]*w2/(w2 - c[9]));
    case OTB_N_SCHOTT:
        return sqrt(c[0] + c[1]*w2 + c[2]/w2 + c[3]/(w2*w2) + c[4]/pow(w2, 3.0) + c[5]/pow(w2, 4.0));
    case OTB_N_HERZBERGER: {
        double L = 1/(w2 - 0.028);
        return c[0] + c[1]*L + c[2]*(L*L) + c[3]*w2 + c[4]*(w2*w2) + c[5]*pow(w2, 3.0);
    }
    case OTB_N_HANDBOOK1: return sqrt(c[0] + c[1]/(w2 - c[2]) - c[3]*w2);
    case OTB_N_HANDBOOK2: return sqrt(c[0] + c[1]*w2/(w2 - c[2]) - c[3]*w2);
    case OTB_N_EXTENDED:
        return sqrt(c[0] + c[1]*w2 + c[2]/w2 + c[3]/(w2*w2) + c[4]/pow(w2, 3.0) + c[5]/pow(w2, 4.0)
                    + c[6]/pow(w2, 5.0) + c[7]/pow(w2, 6.0));
    case OTB_N_EXTENDED2:
        return sqrt(c[0] + c[1]*w2 + c[2]/w2 + c[3]/(w2*w2) + c[4]/pow(w2, 3.0) + c[5]/pow(w2, 4.0)
                    + c[6]*(w2*w2) + c[7]*pow(w2, 3.0));
    case OTB_N_EXTENDED3:
        return sqrt(c[0] + c[1]*w2 + c[2]*(w2*w2) + c[3]/w2 + c[4]/(w2*w2) + c[5]/pow(w2, 3.0)
                    + c[6]*pow(w2, 4.0) + c[7]*pow(w2, 5.0) + c[8]/pow(w2, 6.0));
    case OTB_N_DATA:
        return interp_lr0(aux + M.aux_off, aux + M.aux_off + M.aux_n, M.aux_n, wl);
    case OTB_N_FUNCTION:
        return otb_user_f1(M.func_id, wl);
    default:
        return nan("");
    }
}

// TransmissionSpectrum.__call__ times the incoming float32 weight
// (raytracer.py:380, transmission_spectrum.py:73-84, spectrum.py:81-119).  Returns the new float32 weight.
__device__ inline float filter_apply(const OtbFilter& F, const double* __restrict__ aux, float wlf, float w)
{
    const double wl = (double)wlf;
    double T;
    switch (F.type) {
    case OTB_T_CONSTANT: T = F.c[0]; break;
    case OTB_T_DATA: T = interp_lr0(aux + F.aux_off, aux + F.aux_off + F.aux_n, F.aux_n, wl); break;
    case OTB_T_RECTANGLE: T = (F.c[0] <= wl && wl <= F.c[1]) ? F.c[2] : 0.0; break;
    case OTB_T_GAUSSIAN: {
        // The reference evaluates this branch on the raw float32 wavelengths (spectrum.py:113 uses `wl`,
        // NEP-50 keeps Python-float operands weak), so T is a float32 quantity.  It is computed here in
        // float64 and rounded once, which equals a correctly rounded float32 exp in all but rare ties.
        float d = __fsub_rn(wlf, (float)F.c[1]);
        float q = __fdiv_rn(-__fmul_rn(d, d), (float)F.c[3]);
        float e = (float)exp((double)q);
        float Tf = __fmul_rn((float)F.c[0], e);
        if (F.inverse) Tf = 1.0f - Tf;
        return __fmul_rn(w, Tf);
    }
    case OTB_T_FUNCTION: T = otb_user_f1(F.func_id, wl); break;
    default: T = nan(""); break;
    }
    if (F.inverse) T = 1.0 - T;
    return (float)((double)w*T);
}

// CIE 1931 2-degree observers: 471 samples, 360..830 nm in 1 nm steps (observers.py:8-41).
// Table layout: obs[3*j + c] for wavelength 360 + j.
#define OTB_NOBS 471
__device__ __forceinline__ void observer_xyz(const double* __restrict__ obs, double wl, double& X, double& Y, double& Z)
{
    if (!(wl >= 360.0) || !(wl <= 830.0)) { X = Y = Z = (wl != wl) ? wl : 0.0; return; }
    int j = (int)floor(wl - 360.0);
    if (j >= OTB_NOBS - 1) { j = OTB_NOBS - 1; X = obs[3*j]; Y = obs[3*j + 1]; Z = obs[3*j + 2]; return; }
    double xj = 360.0 + (double)j;
    if (wl == xj) { X = obs[3*j]; Y = obs[3*j + 1]; Z = obs[3*j + 2]; return; }
    double d = wl - xj;
    // slope = (fp[j+1] - fp[j]) / (xp[j+1] - xp[j]) with a knot spacing of exactly 1.0
    X = (obs[3*j + 3] - obs[3*j])*d + obs[3*j];
    Y = (obs[3*j + 4] - obs[3*j + 1])*d + obs[3*j + 1];
    Z = (obs[3*j + 5] - obs[3*j + 2])*d + obs[3*j + 2];
}
