"""Geometry presets: the Arizona eye model (Schwiegerling, Field Guide to Visual and Ophthalmic Optics, SPIE
2004), parameterised by accommodation like presets/geometry.py:54-119 of the reference.  The plot-only
vitreous Volume of the reference is omitted (GUI decoration)."""
import numpy as np

from ..elements import Group, Lens, Aperture, Detector
from ..surfaces import ConicSurface, RingSurface, SphericalSurface
from ..media import RefractionIndex


def arizona_eye(adaptation: float = 0., pupil: float = 5.7, r_det: float = 8, pos: list = None) -> Group:
    pos0 = np.array(pos if pos is not None else [0, 0, 0])
    geom = Group(long_desc="Arizona Eye Model", desc="Eye")
    A = adaptation
    d_Aq = 2.97 - 0.04*A       # aqueous thickness
    d_Lens = 3.767 + 0.04*A    # lens thickness
    n_Cornea = RefractionIndex("Abbe", n=1.377, V=57.1, desc="n_Cornea")
    n_Aqueous = RefractionIndex("Abbe", n=1.337, V=61.3, desc="n_Aqueous")
    n_Lens = RefractionIndex("Abbe", n=1.42+0.00256*A-0.00022*A**2, V=51.9, desc="n_Lens")
    n_Vitreous = RefractionIndex("Abbe", n=1.336, V=61.1, desc="n_Vitreous")

    front = ConicSurface(r=5.45, R=7.8, k=-0.25, long_desc="Cornea Anterior")
    back = ConicSurface(r=5.45, R=6.5, k=-0.25, long_desc="Cornea Posterior")
    L0 = Lens(front, back, d1=0, d2=0.55, pos=pos0+[0, 0, 0], n=n_Cornea, n2=n_Aqueous, desc="Cornea")
    geom.add(L0)

    ap = RingSurface(r=5.45, ri=pupil/2, desc="Pupil")
    geom.add(Aperture(ap, pos=pos0+[0, 0, L0.back.pos[2]+d_Aq-1e-9], desc="Pupil"))

    front = ConicSurface(r=5.1, R=12-0.4*A, k=-7.518749+1.285720*A, long_desc="Lens Anterior")
    back = ConicSurface(r=5.1, R=-5.224557+0.2*A, k=-1.353971-0.431762*A, long_desc="Lens Posterior")
    geom.add(Lens(front, back, d1=0, d2=d_Lens, pos=pos0+[0, 0, d_Aq+0.55], n=n_Lens, n2=n_Vitreous, desc="Lens"))

    geom.add(Detector(SphericalSurface(r=r_det, R=-13.4, desc="Retina"), pos=pos0+[0, 0, 24], desc="Retina"))
    return geom
