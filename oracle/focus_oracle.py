"""CPU oracle for Raytracer.focus_search — TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

numpy restatement of the reference's focus cost functions on stored ray sections:
  raytracer.py:1545-1576   section selection, auxiliary lines pa, sb (via RayStorage.rays_by_mask, ray_storage.py:235-293)
  raytracer.py:1354-1420   __focus_search_cost_function (RMS Spot Size, Irradiance Variance, Image (Center) Sharpness)
  raytracer.py:1422-1447   __focus_rms_spot_direct_solution
Pinned: tests/test_oracle_golden.py::test_focus_oracle_matches_reference against values computed by the reference
itself on the fixture bundles (tests/golden/focus_*.npz, generator tools/gen_golden_focus.py)."""
import numpy as np

METHODS = ["RMS Spot Size", "Irradiance Variance", "Image Sharpness", "Image Center Sharpness"]


def lines(p_list, w_list, z):
    """(pa, sb, w) of the rays that have a stored point behind z; p_list (N, nt, 3), w_list (N, nt)"""
    N, nt, _ = p_list.shape
    pos = np.argmax(z < p_list[:, :, 2], axis=1) - 1
    use = pos != -1
    rp = np.where(use)[0]
    pos = pos[rp]
    p1 = np.where(pos < nt - 1, pos + 1, pos)
    p = p_list[rp, pos]
    s = p_list[rp, p1] - p
    s = s/np.sqrt(s[:, 0]**2 + s[:, 1]**2 + s[:, 2]**2)[:, np.newaxis]
    w = w_list[rp, pos]
    pa = p - s/s[:, 2, np.newaxis]*p[:, 2, np.newaxis]
    sb = s/s[:, 2, np.newaxis]
    return pa, sb, w


def binning_indices_2d(x, y, w, Nx, Ny, extent):
    """misc.py:59-91"""
    s = extent[1] - extent[0], extent[3] - extent[2]
    xi = np.floor(Nx/s[0]*(x - extent[0])).astype(np.int32)
    yi = np.floor(Ny/s[1]*(y - extent[2])).astype(np.int32)
    yi[y == extent[3]] = Ny - 1
    xi[x == extent[1]] = Nx - 1
    outside = (xi < 0) | (yi < 0) | (yi >= Ny) | (xi >= Nx)
    wm = np.where(outside, 0, w)
    yi[outside] = 0
    xi[outside] = 0
    return xi, yi, wm


def cost(z_pos, mode, pa, sb, w):
    """raytracer.py:1354-1420"""
    ph = pa + sb*z_pos
    x, y = ph[:, 0], ph[:, 1]
    if mode == "RMS Spot Size":
        return np.sqrt(np.cov(x, aweights=w) + np.cov(y, aweights=w))
    N_px = 100*int(1 + np.sqrt(w.shape[0])/1500)
    N_px = N_px if N_px % 2 else N_px + 1
    ext = [x.min(), x.max(), y.min(), y.max()]
    xi, yi, wm = binning_indices_2d(x, y, w, N_px, N_px, ext)
    Im = np.zeros((N_px, N_px))
    np.add.at(Im, (yi, xi), wm)
    if mode in ["Image Sharpness", "Image Center Sharpness"]:
        if mode == "Image Center Sharpness":
            Y, X = np.mgrid[-1:1:N_px*1j, -1:1:N_px*1j]
            R = np.sqrt(X**2 + Y**2)
            win = np.where(R > 1, 0, 1 + np.cos(R*np.pi))
            Im0 = Im*win
            if (Im0s := Im0.sum()):
                Im0 *= 1/Im0s
        else:
            Im0 = Im
        return -(((Im0[1:] - Im0[:-1])**2).sum() + ((Im0[:, 1:] - Im0[:, :-1])**2).sum())
    Im = Im[Im > 0]
    Ap = (ext[1] - ext[0])*(ext[3] - ext[2])/N_px**2
    return -np.log(Im.var()/Ap**2)


def rms_direct(pa, sb, w, bounds):
    """raytracer.py:1422-1447: returns (z of the smallest RMS spot, cost there)"""
    pb0 = np.average(pa + sb*bounds[0], axis=0, weights=w)
    pb1 = np.average(pa + sb*bounds[1], axis=0, weights=w)
    vx, vy, vz = pb1[0] - pb0[0], pb1[1] - pb0[1], bounds[1] - bounds[0]
    dx, dy = pa[:, 0] - pb0[0], pa[:, 1] - pb0[1]
    dtx, dty = sb[:, 0] - vx/vz, sb[:, 1] - vy/vz
    w2 = w**2
    dnorm = np.sum(w2*dtx**2 + w2*dty**2)
    d = -np.sum(dtx*dx*w2 + dty*dy*w2)/dnorm if dnorm else np.mean(bounds)
    d = np.clip(d, bounds[0], bounds[1])
    return float(d), float(cost(d, "RMS Spot Size", pa, sb, w))
