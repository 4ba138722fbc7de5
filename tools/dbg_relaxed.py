import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.simplefilter("ignore")
import numpy as np
import optrace_b200 as ot, scenes, golden_util as gu
from optrace_b200.scene import flatten_raytracer
ot.global_options.show_warnings = False
g = gu.load("microscope")
RT = scenes.SCENES["microscope"](ot)
RT.arithmetic = "relaxed"
p0, s0, pol0, w0, wl, hz = gu.bundle(g)
RT.trace_rays(p0, s0, pol0, w0, wl, N_list=g["N_list"])
W, Wr = RT.rays.w_list, g["w_list"]
rel = np.abs(W - Wr)/np.maximum(np.abs(Wr), 1e-30)
bad = np.argwhere(rel > 1e-6)
print("bad entries", bad.shape[0], "first sections", np.unique(bad[:, 1])[:10])
fs = flatten_raytracer(RT)
sec = int(bad[:, 1].min())
st = fs.steps[sec - 1]
print("step", sec - 1, st, fs.surfaces[st["surface"]]["kind"], fs.surfaces[st["surface"]]["par"][:8], "media", fs.media[st["medium_after"]])
r = bad[bad[:, 1] == sec][0, 0]
print("ray", r, "w", W[r, sec-1:sec+1], Wr[r, sec-1:sec+1], "n", RT.rays.n_list[r, sec-1:sec+1], "p", RT.rays.p_list[r, sec], g["p_list"][r, sec])
print("pol", RT.rays.pol_list[r, sec-1:sec+1], g["pol_list"][r, sec-1:sec+1])
print("count bad rays at that section", np.count_nonzero(bad[:, 1] == sec), "of", W.shape[0])
