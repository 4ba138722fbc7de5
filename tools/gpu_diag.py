"""GPU-side diagnostic: per-scene error statistics of the CUDA trace vs the golden fixtures."""
import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
warnings.simplefilter("ignore")
import optrace_b200 as ot
import scenes, golden_util as gu

names = sys.argv[1:] or ["cosine_surfaces", "zoo_numeric"]
for name in names:
    g = gu.load(name)
    RT = scenes.SCENES[name](ot)
    p0, s0, pol0, w0, wl, hz = gu.bundle(g)
    RT.trace_rays(p0, s0, pol0, w0, wl, hurb_z=hz, N_list=g["N_list"])
    R = RT.rays
    print("==", name, "msgs equal", np.array_equal(RT._msgs, g["msgs"]))
    P, Pr = R.p_list, g["p_list"]
    d = np.abs(P - Pr)
    for i in range(P.shape[1]):
        di = d[:, i].max(axis=1)
        bad = np.nonzero(di > 1e-12)[0]
        print(f" section {i}: max abs err {di.max():.3e}, rays >1e-12: {bad.size}, >1e-9: {np.count_nonzero(di > 1e-9)}")
        for r in bad[:3]:
            print("    ray", r, "gpu", P[r, i], "ref", Pr[r, i], "w", R.w_list[r, i], g["w_list"][r, i])
    print(" s maxabs", np.nanmax(np.abs(R.s0_list - g["s_list"])), " w maxrel", gu.maxrel(R.w_list, g["w_list"]))
    if "pol_list" in g:
        print(" pol maxabs", np.nanmax(np.abs(R.pol_list.astype(float) - g["pol_list"])))
