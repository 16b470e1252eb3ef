"""Golden-vector case list shared by make_golden.py (runs the reference on a B200) and the parity tests.

Inputs are regenerated from the seed (`make_input`), so the .npz files only hold the reference's outputs.
Sizes cover: even / odd dims, both parities of hlen/2 (inverse centring branches, separable.cu:252-264),
hlen from 2 to 40, 1-D batched, SWT, non-separable, Haar special case, level clamping.
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# name, shape, wname, levels, do_separable, do_swt, ndim
CASES = [
    ("c1_haar1d_4096",      (1, 4096),  "haar",    1, 1, 0, 1),   # BASELINE.json configs[0]
    ("dwt2_db7_128",        (128, 128), "db7",     3, 1, 0, 2),   # the headline wavelet, h2 odd
    ("dwt2_lena_db7",       "lena",     "db7",     3, 1, 0, 2),
    ("dwt2_sym8_64",        (64, 64),   "sym8",    2, 1, 0, 2),   # h2 even
    ("dwt2_db2_57x64",      (57, 64),   "db2",     3, 1, 0, 2),   # odd rows
    ("dwt2_db3_33x47",      (33, 47),   "db3",     2, 1, 0, 2),   # odd rows and cols
    ("dwt2_bior22_48x40",   (48, 40),   "bior2.2", 2, 1, 0, 2),
    ("dwt2_rbio31_40x36",   (40, 36),   "rbio3.1", 3, 1, 0, 2),
    ("dwt2_coif2_72x88",    (72, 88),   "coif2",   2, 1, 0, 2),
    ("dwt2_db10_96",        (96, 96),   "db10",    2, 1, 0, 2),
    ("dwt2_db20_160",       (160, 160), "db20",    2, 1, 0, 2),   # hlen 40
    ("dwt2_clamp_db7_40",   (40, 40),   "db7",     5, 1, 0, 2),   # levels clamped to ilog2(40/13)=1
    ("haar2_64",            (64, 64),   "haar",    3, 1, 0, 2),
    ("haar2_33x47",         (33, 47),   "haar",    3, 1, 0, 2),
    ("haar2_even_lv_50x70", (50, 70),   "db1",     2, 1, 0, 2),   # alias name, even level count (D2D fix-up path)
    ("haar1_3x1001",        (3, 1001),  "haar",    4, 1, 0, 1),
    ("dwt1_db7_4x512",      (4, 512),   "db7",     3, 1, 0, 1),
    ("dwt1_sym4_5x333",     (5, 333),   "sym4",    4, 1, 0, 1),
    ("dwt1_db2_1x64",       (1, 64),    "db2",     2, 1, 0, 1),
    ("swt2_sym8_64",        (64, 64),   "sym8",    2, 1, 1, 2),
    ("swt2_db2_48x56",      (48, 56),   "db2",     3, 1, 1, 2),
    ("swt2_haar_32",        (32, 32),   "haar",    3, 1, 1, 2),
    ("swt2_db3_33x47",      (33, 47),   "db3",     2, 1, 1, 2),
    ("swt1_db4_3x256",      (3, 256),   "db4",     3, 1, 1, 1),
    ("swt1_haar_2x101",     (2, 101),   "haar",    2, 1, 1, 1),
    ("ns2_db7_64",          (64, 64),   "db7",     2, 0, 0, 2),
    ("ns2_db2_33x47",       (33, 47),   "db2",     2, 0, 0, 2),
    ("ns2_coif2_56x48",     (56, 48),   "coif2",   2, 0, 0, 2),
    ("ns2_sym4_3lv_64x72",  (64, 72),   "sym4",    3, 0, 0, 2),
    ("ns2_haar_32",         (32, 32),   "haar",    2, 0, 0, 2),
    ("nsswt2_db2_32",       (32, 32),   "db2",     2, 0, 1, 2),
    ("nsswt2_db3_24x40",    (24, 40),   "db3",     3, 0, 1, 2),
]

# cases that also carry the four thresholded variants (keeps the committed fixtures small)
THRESH_CASES = {"c1_haar1d_4096", "dwt2_db7_128", "dwt2_db3_33x47", "haar2_33x47", "dwt1_sym4_5x333", "swt2_db2_48x56",
                "swt1_haar_2x101", "ns2_db2_33x47", "nsswt2_db2_32"}

THRESH = [  # (tag, kind, beta, do_thresh_appcoeffs, normalize)
    ("soft", "soft", 10.0, 0, 0),
    ("softan", "soft", 25.0, 1, 1),
    ("hard", "hard", 10.0, 0, 0),
    ("hardan", "hard", 25.0, 1, 1),
]


def make_input(name, shape):
    """float32 input of SURVEY section 8(d): N(128, 50^2) from default_rng(seed) -- seed = crc of the case name."""
    if isinstance(shape, str):
        return np.load(os.path.join(HERE, "lena_crop128.npy"))
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    return (rng.standard_normal(shape) * 50 + 128).astype(np.float32)
