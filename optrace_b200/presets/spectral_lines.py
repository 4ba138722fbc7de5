"""Fraunhofer / spectral lines in nm (physical constants; reference: presets/spectral_lines.py)."""
h, g, F_, F, e, d, D, C_, C, r, A_ = (404.6561, 435.8343, 479.9914, 486.1327, 546.0740, 587.5618, 589.2938,
                                      643.8469, 656.272, 706.5188, 768.2)
all_lines = [h, g, F_, F, e, d, D, C_, C, r, A_]
FDC, FdC, FeC, F_eC_ = [F, D, C], [F, d, C], [F, e, C], [F_, e, C_]
rgb = [464.3118, 549.1321, 611.2826]
all_line_combinations = [FDC, FdC, FeC, F_eC_, rgb]
