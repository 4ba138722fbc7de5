"""Golden vectors for Raytracer.focus_search, generated with the REFERENCE itself.

The frozen bundle of a fixture scene (tests/golden/<scene>.npz) is injected into the reference's Raytracer; the
reference's own cost function (Raytracer.__focus_search_cost_function) is evaluated at fixed z positions for all
four methods and its deterministic RMS solution (focus_search("RMS Spot Size")) is recorded.
Build container only (needs /root/reference)."""
import pathlib
import sys
import warnings

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "tests"))
from refharness import import_reference  # noqa: E402
import golden_util as gu  # noqa: E402
import scenes  # noqa: E402

CASES = {"spherical_aberration": 23.0, "double_gauss": 120.0}     # scene -> z_start


def main():
    ot = import_reference()
    from optrace.tracer.geometry import RaySource
    ot.global_options.multithreading = False
    ot.global_options.show_progress_bar = False
    warnings.simplefilter("ignore")
    for scene, z_start in CASES.items():
        g = gu.load(scene)
        RT = scenes.SCENES[scene](ot)
        B = np.concatenate(([0], np.cumsum(g["N_list"])))
        calls = [0]
        orig = RaySource.create_rays

        def inject(self, N_, no_pol=False, power=None):
            k = calls[0]
            calls[0] += 1
            a, b = B[k], B[k + 1]
            pol = np.broadcast_to(np.nan, (b - a, 3)) if no_pol else g["pol0"][a:b].astype(np.float64)
            return g["p0"][a:b].copy(), g["s0"][a:b].copy(), pol, g["w0"][a:b].copy(), g["wl"][a:b].astype(np.float64)

        RaySource.create_rays = inject
        try:
            np.random.seed(99)
            RT.trace(int(g["N"]))
        finally:
            RaySource.create_rays = orig
        assert np.array_equal(RT.rays.p_list, g["p_list"]) and np.array_equal(RT.rays.w_list, g["w_list"])

        res, info = RT.focus_search("RMS Spot Size", z_start)
        bounds = info["bounds"]
        # the prelude of focus_search (raytracer.py:1545-1576) with the reference's own calls
        rays_pos = np.ones(RT.rays.N, dtype=bool)
        z = bounds[0] + RT.N_EPS
        pos = np.argmax(z < RT.rays.p_list[:, :, 2], axis=1) - 1
        rays_pos[pos == -1] = False
        rp = np.where(rays_pos)[0]
        p, s, _, w, _, _, _ = RT.rays.rays_by_mask(rays_pos, pos[rp], ret=[1, 1, 0, 1, 0, 0, 0])
        pa = p - s/s[:, 2, np.newaxis]*p[:, 2, np.newaxis]
        sb = s/s[:, 2, np.newaxis]
        zs = np.linspace(bounds[0], bounds[1], 9)[1:-1]
        cost = RT._Raytracer__focus_search_cost_function
        out = dict(z_start=z_start, bounds=np.array(bounds), N_use=info["N"], rms_x=res.x, rms_fun=res.fun,
                   rms_pos=np.array(info["pos"]), zs=zs, methods=np.array(RT.focus_search_methods))
        for k, m in enumerate(RT.focus_search_methods):
            out[f"cost{k}"] = np.array([cost(float(zz), m, pa, sb, w) for zz in zs])
        path = ROOT / "tests" / "golden" / f"focus_{scene}.npz"
        np.savez_compressed(path, **out)
        print(scene, "bounds", bounds, "N_use", info["N"], "rms focus", res.x, "->", path.name)


if __name__ == "__main__":
    main()
