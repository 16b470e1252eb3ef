"""CPU oracle for the PDWT hot path -- TEST INFRASTRUCTURE ONLY.

`oracle.Wavelets` mirrors the reference's `Wavelets` class (wt.h:20-76, wt.cu:84-508) on numpy arrays by
calling the C restatement in `pdwt_oracle.c` (one function per reference kernel / driver, each citing the
file:line it follows).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this package; nothing under pdwt_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

W_INIT, W_FORWARD, W_INVERSE, W_THRESHOLD, W_CREATION_ERROR = 0, 1, 2, 3, 4

_fp = C.POINTER(C.c_float)
_fpp = C.POINTER(_fp)


class OrcInfo(C.Structure):  # struct w_info, utils.h:9-19
    _fields_ = [(n, C.c_int) for n in ("ndims", "Nr", "Nc", "nlevels", "do_swt", "hlen")]


class OrcFilters(C.Structure):
    _fields_ = [("hlen", C.c_int), ("L", C.c_float * 40), ("H", C.c_float * 40), ("IL", C.c_float * 40),
                ("IH", C.c_float * 40)]


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (oracle/Makefile).  Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pdwt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        drv = [_fp, _fpp, _fp, OrcInfo, C.POINTER(OrcFilters)]
        for name in ("orc_forward_separable_2d", "orc_forward_separable_1d", "orc_inverse_separable_2d",
                     "orc_inverse_separable_1d", "orc_forward_swt_separable_2d", "orc_forward_swt_separable_1d",
                     "orc_inverse_swt_separable_2d", "orc_inverse_swt_separable_1d", "orc_forward_nonseparable_2d",
                     "orc_inverse_nonseparable_2d", "orc_forward_swt_nonseparable_2d",
                     "orc_inverse_swt_nonseparable_2d"):
            getattr(L, name).argtypes = drv
            getattr(L, name).restype = C.c_int
        for name in ("orc_haar_forward_2d", "orc_haar_inverse_2d", "orc_haar_forward_1d", "orc_haar_inverse_1d"):
            getattr(L, name).argtypes = drv[:4]
            getattr(L, name).restype = C.c_int
        L.orc_filters_lookup.argtypes = [C.c_char_p, C.c_int, C.POINTER(OrcFilters)]
        L.orc_filters_lookup.restype = C.c_int
        L.orc_filters_custom.argtypes = [C.POINTER(OrcFilters), C.c_int, _fp, _fp, _fp, _fp]
        L.orc_threshold.argtypes = [_fpp, C.c_float, OrcInfo, C.c_int, C.c_int, C.c_int]
        L.orc_threshold.restype = None
        L.orc_proj_linf.argtypes = [_fpp, C.c_float, OrcInfo, C.c_int]
        L.orc_proj_linf.restype = None
        L.orc_shrink.argtypes = [_fpp, C.c_float, OrcInfo, C.c_int]
        L.orc_shrink.restype = None
        L.orc_group_soft_thresh.argtypes = [_fpp, C.c_float, OrcInfo, C.c_int, C.c_int, C.c_int]
        L.orc_group_soft_thresh.restype = None
        L.orc_add_coeffs.argtypes = [_fpp, _fpp, OrcInfo, C.c_float]
        L.orc_add_coeffs.restype = None
        L.orc_circshift.argtypes = [_fp, _fp, OrcInfo, C.c_int, C.c_int]
        L.orc_circshift.restype = None
        L.orc_norm1.argtypes = [_fpp, OrcInfo]
        L.orc_norm1.restype = C.c_float
        L.orc_norm2sq.argtypes = [_fpp, OrcInfo, C.c_int]
        L.orc_norm2sq.restype = C.c_float
        L.orc_ilog2.argtypes = [C.c_int]
        L.orc_ilog2.restype = C.c_int
        _LIB = L
    return _LIB


def div2(n: int) -> int:  # w_div2, utils.cu:24-27
    return (n + 1) // 2


def filters(wname: str, do_swt: int = 0):
    """(hlen, L, H, IL, IH) as float32 arrays; hlen == 2 with empty taps for the Haar aliases when not SWT."""
    f = OrcFilters()
    hlen = lib().orc_filters_lookup(wname.encode(), int(do_swt), C.byref(f))
    if hlen < 0:
        raise KeyError(wname)
    g = lambda a: np.array(a[:hlen], dtype=np.float32)
    return hlen, g(f.L), g(f.H), g(f.IL), g(f.IH)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_fp)


class Wavelets:
    """numpy mirror of the reference class (same constructor arguments, methods and state rules)."""

    def __init__(self, img, wname: str, levels: int, do_separable: int = 1, do_swt: int = 0, ndim: int = 2):
        img = np.ascontiguousarray(img, dtype=np.float32)
        if img.ndim == 1:
            img = img[None, :]
        Nr, Nc = img.shape
        self.L = lib()
        self.wname = wname
        self.do_separable = int(do_separable)
        self.state = W_INIT
        self.image = img.copy()
        if Nr == 1:
            ndim = 1  # wt.cu:133-136
        if ndim == 1:
            self.do_separable = 1  # wt.cu:138-142
        self.filt = OrcFilters()
        hlen = self.L.orc_filters_lookup(wname.encode(), int(do_swt), C.byref(self.filt))
        levels = max(int(levels), 1)  # wt.cu:111-114
        if hlen <= 0:  # the reference hangs here; the restatement reports the error state
            self.state = W_CREATION_ERROR
            self.info = OrcInfo(ndim, Nr, Nc, levels, int(do_swt), 0)
            return
        N = min(Nr, Nc) if ndim == 2 else Nc
        wmaxlev = self.L.orc_ilog2(N // (hlen - 1))  # wt.cu:156-165
        levels = min(levels, wmaxlev)
        self.info = OrcInfo(ndim, Nr, Nc, levels, int(do_swt), hlen)
        self.tmp = np.zeros(2 * Nr * Nc, dtype=np.float32)  # wt.cu:128-130
        # w_create_coeffs_buffer(_1d), common.cu:400-445
        self.coeffs = []
        nr, nc = Nr, Nc
        per = 3 if ndim == 2 else 1
        for _ in range(levels):
            if not do_swt:
                if ndim == 2:
                    nr = div2(nr)
                nc = div2(nc)
            for _ in range(per):
                self.coeffs.append(np.zeros((nr, nc), dtype=np.float32))
        self.app_shape = (nr, nc)  # true A_L size
        n0r = Nr if (do_swt or ndim == 1) else div2(Nr)
        n0c = Nc if do_swt else div2(Nc)
        self.coeffs.insert(0, np.zeros(n0r * n0c, dtype=np.float32))  # level-1 sized scratch + A_L
        self._cptr = (_fp * len(self.coeffs))(*[_ptr(c) for c in self.coeffs])

    # ---- dispatch, wt.cu:236-307 --------------------------------------------------------------------
    def _driver(self, direction: str):
        w = self.info
        haar = (w.hlen == 2 and not w.do_swt)
        d = "1d" if w.ndims == 1 else "2d"
        if haar:
            return getattr(self.L, f"orc_haar_{direction}_{d}"), False
        swt = "swt_" if w.do_swt else ""
        sep = "separable" if (self.do_separable or w.ndims == 1) else "nonseparable"
        return getattr(self.L, f"orc_{direction}_{swt}{sep}_{d}"), True

    def forward(self):
        if self.state == W_CREATION_ERROR:
            return
        fn, needs_f = self._driver("forward")
        args = [_ptr(self.image), self._cptr, _ptr(self.tmp), self.info]
        if needs_f:
            args.append(C.byref(self.filt))
        fn(*args)
        self.state = W_FORWARD

    def inverse(self):
        if self.state in (W_INVERSE, W_CREATION_ERROR):
            return
        fn, needs_f = self._driver("inverse")
        args = [_ptr(self.image), self._cptr, _ptr(self.tmp), self.info]
        if needs_f:
            args.append(C.byref(self.filt))
        fn(*args)
        self.state = W_INVERSE

    def soft_threshold(self, beta, do_thresh_appcoeffs=0, normalize=0):
        if self.state == W_INVERSE:
            return
        self.L.orc_threshold(self._cptr, float(beta), self.info, int(do_thresh_appcoeffs), int(normalize), 0)

    def hard_threshold(self, beta, do_thresh_appcoeffs=0, normalize=0):
        if self.state == W_INVERSE:
            return
        self.L.orc_threshold(self._cptr, float(beta), self.info, int(do_thresh_appcoeffs), int(normalize), 1)

    # ---- the other proximal operators and helpers, wt.cu:330-368, 624-657 ----------------------------------
    def group_soft_threshold(self, beta, do_thresh_appcoeffs=0, normalize=0, variant=0):
        if self.state == W_INVERSE:
            return
        self.L.orc_group_soft_thresh(self._cptr, float(beta), self.info, int(do_thresh_appcoeffs), int(normalize),
                                     int(variant))

    def shrink(self, beta, do_thresh_appcoeffs=1):
        if self.state == W_INVERSE:
            return
        self.L.orc_shrink(self._cptr, float(beta), self.info, int(do_thresh_appcoeffs))

    def proj_linf(self, beta, do_thresh_appcoeffs=1):
        if self.state == W_INVERSE:
            return
        self.L.orc_proj_linf(self._cptr, float(beta), self.info, int(do_thresh_appcoeffs))

    def circshift(self, sr, sc, inplace=1):
        out = np.empty_like(self.image)
        self.L.orc_circshift(_ptr(self.image), _ptr(out), self.info, int(sr), int(sc))
        if inplace:
            self.image[...] = out
        else:
            self.tmp[: out.size] = out.ravel()

    def add_wavelet(self, W, alpha=1.0) -> int:
        if self.info.nlevels != W.info.nlevels or self.wname.lower() != W.wname.lower():
            return -1
        if self.state == W_INVERSE or W.state == W_INVERSE:
            return 1
        if (self.info.Nr, self.info.Nc, self.info.ndims) != (W.info.Nr, W.info.Nc, W.info.ndims):
            return -2
        if bool(self.info.do_swt) != bool(W.info.do_swt):
            return -3
        self.L.orc_add_coeffs(self._cptr, W._cptr, self.info, float(alpha))
        return 0

    def norm1(self) -> float:
        return float(self.L.orc_norm1(self._cptr, self.info))

    def norm2sq(self, ref_1d_bug: int = 0) -> float:
        return float(self.L.orc_norm2sq(self._cptr, self.info, int(ref_1d_bug)))

    # ---- accessors, wt.cu:421-508 ---------------------------------------------------------------------
    def get_image(self) -> np.ndarray:
        return self.image.copy()

    def set_image(self, img):
        self.image[...] = np.asarray(img, dtype=np.float32).reshape(self.image.shape)
        self.state = W_INIT

    def coeff_shape(self, num: int):
        return self.app_shape if num == 0 else self.coeffs[num].shape

    def get_coeff(self, num: int) -> np.ndarray:
        if self.state == W_INVERSE:
            return None
        if num == 0:
            nr, nc = self.app_shape
            return self.coeffs[0][: nr * nc].reshape(nr, nc).copy()
        return self.coeffs[num].copy()

    def set_coeff(self, arr, num: int):
        arr = np.asarray(arr, dtype=np.float32)
        if num == 0:
            self.coeffs[0][: arr.size] = arr.ravel()
        else:
            self.coeffs[num][...] = arr.reshape(self.coeffs[num].shape)

    @property
    def ncoeffs(self) -> int:
        return len(self.coeffs)
