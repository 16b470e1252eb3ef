"""Golden-vector cases for the proximal operators / helpers of the class: group soft threshold,
proj_linf, shrink, add_wavelet, circshift.  Shared by make_golden_prox.py (runs the reference on a B200) and the tests."""
import numpy as np

# name, shape, wname, levels, do_separable, do_swt, ndim
PROX_CASES = [
    ("prox_dwt2_db3_33x47", (33, 47), "db3", 2, 1, 0, 2),     # odd sizes
    ("prox_dwt2_db7_64x96", (64, 96), "db7", 2, 1, 0, 2),
    ("prox_swt2_sym4_32x40", (32, 40), "sym4", 2, 1, 1, 2),   # A at full size
    ("prox_dwt1_db4_5x333", (5, 333), "db4", 3, 1, 0, 1),     # 1-D, odd halves
    ("prox_haar2_48", (48, 48), "haar", 3, 1, 0, 2),
]

# (tag, method, args)
PROX_OPS = [
    ("gs", "group_soft_threshold", (30.0, 0, 0)),
    ("gsan", "group_soft_threshold", (60.0, 1, 1)),
    ("shr", "shrink", (0.7, 1)),
    ("shr0", "shrink", (0.25, 0)),
    ("linf", "proj_linf", (12.5, 1)),
    ("linf0", "proj_linf", (40.0, 0)),
]
ADD_ALPHA = 0.75
SHIFTS = [(5, -7), (-3, 11), (0, 1)]


def prox_input(shape, seed):
    return (np.random.default_rng(seed).standard_normal(shape) * 50 + 128).astype(np.float32)
