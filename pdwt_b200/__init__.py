"""pdwt_b200 -- B200-native (sm_100a) implementation of PDWT's hot path behind PDWT's own `Wavelets` interface.

This module is the host-side mirror of the reference class for Python callers (what the external `pypwt` Cython
binding is to libpdwt.so): a thin ctypes layer over the C ABI of `libpdwt_b200.so` (include/pdwt_b200.h).  Same
constructor arguments, method names, state rules and error behaviour as reference `src/wt.h:20-76` /
`src/wt.cu:84-508`.  There is NO CPU fallback: if the CUDA library is missing or no device is usable, construction
raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["Wavelets", "lib", "build", "PdwtError", "W_INIT", "W_FORWARD", "W_INVERSE", "W_THRESHOLD",
           "W_CREATION_ERROR", "wavelet_names", "filters"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libpdwt_b200.so")
_lib = None

# enum w_state, wt.h:8-17
W_INIT, W_FORWARD, W_INVERSE, W_THRESHOLD, W_CREATION_ERROR, W_FORWARD_ERROR, W_INVERSE_ERROR, W_THRESHOLD_ERROR = range(8)

PDWT_OK, PDWT_ERR_ARG, PDWT_ERR_WAVELET, PDWT_ERR_FILTER_LEN, PDWT_ERR_STATE, PDWT_ERR_ALLOC, PDWT_ERR_CUDA = \
    0, -1, -2, -3, -4, -5, -6

_fp = C.POINTER(C.c_float)
_fpp = C.POINTER(_fp)


class PdwtError(RuntimeError):
    pass


class WInfo(C.Structure):  # struct w_info, utils.h:9-19
    _fields_ = [(n, C.c_int) for n in ("ndims", "Nr", "Nc", "nlevels", "do_swt", "hlen")]


class ProfileEntry(C.Structure):  # pdwt_profile_entry, include/pdwt_b200.h
    _fields_ = [("name", C.c_char * 64), ("launches", C.c_int), ("ms_total", C.c_double), ("ms_min", C.c_double),
                ("ms_max", C.c_double)]


def build(force: bool = False) -> str:
    from .build import build as _b
    return _b(force=force)


_DRIVERS = ["pdwt_forward_separable", "pdwt_forward_separable_1d", "pdwt_inverse_separable",
            "pdwt_inverse_separable_1d", "pdwt_forward_swt_separable", "pdwt_forward_swt_separable_1d",
            "pdwt_inverse_swt_separable", "pdwt_inverse_swt_separable_1d", "pdwt_haar_forward2d",
            "pdwt_haar_inverse2d", "pdwt_haar_forward1d", "pdwt_haar_inverse1d", "pdwt_forward_nonseparable",
            "pdwt_inverse_nonseparable", "pdwt_forward_swt_nonseparable", "pdwt_inverse_swt_nonseparable"]


def lib() -> C.CDLL:
    """Load libpdwt_b200.so (built in-tree by pdwt_b200/build.py) and declare the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise PdwtError(f"{_LIBPATH} is missing: run `python -m pdwt_b200.build` (there is no CPU fallback)")
    L = C.CDLL(_LIBPATH)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.pdwt_version.restype = C.c_char_p
    L.pdwt_last_cuda_error_string.restype = C.c_char_p
    L.pdwt_wavelet_name.restype = C.c_char_p
    L.pdwt_wavelet_name.argtypes = [ci]
    L.pdwt_launch_count.restype = C.c_longlong
    L.pdwt_filters_create.argtypes = [C.POINTER(vp), C.c_char_p, ci]
    L.pdwt_filters_create_custom.argtypes = [C.POINTER(vp), ci, _fp, _fp, _fp, _fp]
    L.pdwt_filters_destroy.argtypes = [vp]
    L.pdwt_filters_destroy.restype = None
    L.pdwt_filters_hlen.argtypes = [vp]
    L.pdwt_filters_get.argtypes = [vp, _fp, _fp, _fp, _fp]
    drv = [vp, vp, C.POINTER(vp), vp, WInfo, ci, vp]
    for n in _DRIVERS:
        getattr(L, n).argtypes = drv
    L.pdwt_forward.argtypes = drv + [ci]
    L.pdwt_inverse.argtypes = drv + [ci]
    for n in ("pdwt_call_soft_thresh", "pdwt_call_hard_thresh"):
        getattr(L, n).argtypes = [C.POINTER(vp), cf, WInfo, ci, ci, ci, vp]
    for n in ("pdwt_norm1", "pdwt_norm2sq"):
        getattr(L, n).argtypes = [C.POINTER(vp), WInfo, ci, _fp, vp]
    L.pdwt_div2.argtypes = [ci]
    L.pdwt_num_coeffs.argtypes = [WInfo]
    L.pdwt_coeff_dims.argtypes = [WInfo, ci, C.POINTER(ci), C.POINTER(ci)]
    L.pdwt_coeff_alloc_elems.argtypes = [WInfo, ci]
    L.pdwt_coeff_alloc_elems.restype = C.c_size_t
    L.pdwt_max_level.argtypes = [ci, ci, ci, ci]
    L.pdwt_wavelets_create.argtypes = [C.POINTER(vp), vp, ci, ci, C.c_char_p, ci, ci, ci, ci, ci, ci, ci]
    L.pdwt_wavelets_copy.argtypes = [C.POINTER(vp), vp]
    L.pdwt_wavelets_destroy.argtypes = [vp]
    L.pdwt_wavelets_destroy.restype = None
    for n in ("pdwt_wavelets_forward", "pdwt_wavelets_inverse", "pdwt_wavelets_sync", "pdwt_wavelets_state",
              "pdwt_wavelets_batch", "pdwt_wavelets_do_separable"):
        getattr(L, n).argtypes = [vp]
    for n in ("pdwt_wavelets_soft_threshold", "pdwt_wavelets_hard_threshold"):
        getattr(L, n).argtypes = [vp, cf, ci, ci]
    L.pdwt_wavelets_group_soft_threshold.argtypes = [vp, cf, ci, ci]
    L.pdwt_wavelets_shrink.argtypes = [vp, cf, ci]
    L.pdwt_wavelets_proj_linf.argtypes = [vp, cf, ci]
    L.pdwt_wavelets_circshift.argtypes = [vp, ci, ci, ci]
    L.pdwt_wavelets_add_wavelet.argtypes = [vp, vp, cf]
    L.pdwt_wavelets_current_shift.argtypes = [vp, C.POINTER(ci), C.POINTER(ci)]
    L.pdwt_call_group_soft_thresh.argtypes = [C.POINTER(vp), cf, WInfo, ci, ci, ci, vp]
    for n in ("pdwt_call_proj_linf", "pdwt_shrink"):
        getattr(L, n).argtypes = [C.POINTER(vp), cf, WInfo, ci, ci, vp]
    L.pdwt_add_coeffs.argtypes = [C.POINTER(vp), C.POINTER(vp), WInfo, cf, ci, vp]
    L.pdwt_call_circshift.argtypes = [vp, vp, WInfo, ci, ci, ci, ci, vp]
    for n in ("pdwt_wavelets_norm1", "pdwt_wavelets_norm2sq", "pdwt_wavelets_get_image"):
        getattr(L, n).argtypes = [vp, vp]
    L.pdwt_wavelets_get_coeff.argtypes = [vp, vp, ci]
    L.pdwt_wavelets_set_image.argtypes = [vp, vp, ci]
    L.pdwt_wavelets_set_coeff.argtypes = [vp, vp, ci, ci]
    L.pdwt_wavelets_set_filters_forward.argtypes = [vp, C.c_char_p, C.c_uint, _fp, _fp]
    L.pdwt_wavelets_set_filters_inverse.argtypes = [vp, _fp, _fp]
    L.pdwt_wavelets_set_filters_forward_2d.argtypes = [vp, C.c_char_p, C.c_uint, _fp, _fp, _fp, _fp]
    L.pdwt_wavelets_set_filters_inverse_2d.argtypes = [vp, _fp, _fp, _fp, _fp]
    L.pdwt_filters_set_2d.argtypes = [vp, ci, _fp, _fp, _fp, _fp]
    L.pdwt_filters_has_2d.argtypes = [vp, ci]
    L.pdwt_wavelets_set_stream.argtypes = [vp, vp]
    L.pdwt_wavelets_set_async.argtypes = [vp, ci]
    L.pdwt_wavelets_invalidate_norm_cache.argtypes = [vp]
    L.pdwt_wavelets_info.argtypes = [vp]
    L.pdwt_wavelets_info.restype = WInfo
    L.pdwt_wavelets_wname.argtypes = [vp]
    L.pdwt_wavelets_wname.restype = C.c_char_p
    for n in ("pdwt_wavelets_image_int_ptr", "pdwt_wavelets_tmp_int_ptr"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = C.c_ssize_t
    L.pdwt_wavelets_coeff_int_ptr.argtypes = [vp, ci]
    L.pdwt_wavelets_coeff_int_ptr.restype = C.c_ssize_t
    L.pdwt_wavelets_launch_count.argtypes = [vp]
    L.pdwt_wavelets_launch_count.restype = C.c_longlong
    L.pdwt_profile_end.argtypes = [C.POINTER(ProfileEntry), ci]
    _lib = L
    return L


def _check(rc: int, what: str):
    if rc < 0:
        L = lib()
        extra = ""
        if rc == PDWT_ERR_CUDA:
            extra = f" (CUDA error {L.pdwt_last_cuda_error()}: {L.pdwt_last_cuda_error_string().decode()})"
        raise PdwtError(f"{what} failed with pdwt_status {rc}{extra}")
    return rc


def wavelet_names():
    L = lib()
    return [L.pdwt_wavelet_name(i).decode() for i in range(L.pdwt_wavelet_count())]


def filters(wname: str, do_swt: int = 0):
    """(hlen, dec_lo, dec_hi, rec_lo, rec_hi) of a built-in bank -- host only, no device needed."""
    L = lib()
    h = C.c_void_p()
    hlen = L.pdwt_filters_create(C.byref(h), wname.encode(), int(do_swt))
    if hlen < 0:
        raise KeyError(wname)
    arrs = [np.zeros(hlen, dtype=np.float32) for _ in range(4)]
    L.pdwt_filters_get(h, *[a.ctypes.data_as(_fp) for a in arrs])
    L.pdwt_filters_destroy(h)
    return (hlen, *arrs)


def _dev_ptr(x):
    """device pointer of a torch CUDA tensor / anything with data_ptr(), or an int"""
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


class Wavelets:
    """Python face of the `Wavelets` object (reference wt.h:20-76).

    `img` is a float32 numpy array of shape (Nr, Nc) -- or (batch, Nr, Nc) with `batch` planes, the one extension --
    uploaded on construction like the reference does (wt.cu:117-126); a torch CUDA tensor is taken as device memory
    (`memisonhost=0`).  Methods and state rules follow the reference: `inverse()` twice is refused, thresholds and
    `get_coeff` are refused after `inverse()`, `set_image` resets the state.
    """

    def __init__(self, img, wname: str, levels: int, do_separable: int = 1, do_cycle_spinning: int = 0,
                 do_swt: int = 0, ndim: int = 2, shape=None, batch: int | None = None):
        L = lib()
        if L.pdwt_device_count() < 1:
            raise PdwtError("no CUDA device available: pdwt_b200 has no CPU fallback")
        on_device = hasattr(img, "data_ptr")
        if img is None:
            if shape is None:
                raise ValueError("shape is required when img is None")
            shp = tuple(shape)
        else:
            shp = tuple(img.shape)
        if len(shp) == 1:
            shp = (1, shp[0])
        if batch is None:
            batch = shp[0] if len(shp) == 3 else 1
        elif len(shp) == 3 and shp[0] != batch:
            raise ValueError("batch does not match img.shape[0]")
        Nr, Nc = shp[-2], shp[-1]
        self._keep = None
        if img is None:
            ptr = None
        elif on_device:
            if str(img.dtype) != "torch.float32" or not img.is_contiguous():
                raise ValueError("device images must be contiguous float32")
            ptr = C.c_void_p(_dev_ptr(img))
        else:
            self._keep = np.ascontiguousarray(img, dtype=np.float32)
            ptr = self._keep.ctypes.data_as(C.c_void_p)
        h = C.c_void_p()
        _check(L.pdwt_wavelets_create(C.byref(h), ptr, Nr, Nc, wname.encode(), int(levels), 0 if on_device else 1,
                                      int(do_separable), int(do_cycle_spinning), int(do_swt), int(ndim), int(batch)),
               "pdwt_wavelets_create")
        self._h = h
        self._L = L
        self._keep = None
        self.batch = int(batch)
        if self.state == W_CREATION_ERROR and L.pdwt_last_cuda_error() and self.info.hlen > 0 and self.info.nlevels > 0:
            raise PdwtError("Wavelets construction failed on the device: " + L.pdwt_last_cuda_error_string().decode())

    # ---- lifetime -------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.pdwt_wavelets_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def copy(self) -> "Wavelets":
        o = object.__new__(Wavelets)
        h = C.c_void_p()
        _check(self._L.pdwt_wavelets_copy(C.byref(h), self._h), "pdwt_wavelets_copy")
        o._h, o._L, o._keep, o.batch = h, self._L, None, self.batch
        return o

    # ---- public data members of the reference class ----------------------------------------------------------
    @property
    def state(self) -> int:
        return self._L.pdwt_wavelets_state(self._h)

    @property
    def info(self) -> WInfo:
        return self._L.pdwt_wavelets_info(self._h)

    @property
    def wname(self) -> str:
        return self._L.pdwt_wavelets_wname(self._h).decode()

    @property
    def do_separable(self) -> int:
        return self._L.pdwt_wavelets_do_separable(self._h)

    @property
    def nlevels(self) -> int:
        return self.info.nlevels

    @property
    def ncoeffs(self) -> int:
        return self._L.pdwt_num_coeffs(self.info)

    @property
    def launch_count(self) -> int:
        return self._L.pdwt_wavelets_launch_count(self._h)

    def image_int_ptr(self) -> int:
        return self._L.pdwt_wavelets_image_int_ptr(self._h)

    def coeff_int_ptr(self, num: int) -> int:
        """raw device pointer of sub-band `num` (wt.cu:665).  norm1()/norm2sq() right after a threshold return sums the
        threshold kernel produced on its way: after WRITING through this pointer call invalidate_norm_cache()."""
        return self._L.pdwt_wavelets_coeff_int_ptr(self._h, num)

    def invalidate_norm_cache(self):
        _check(self._L.pdwt_wavelets_invalidate_norm_cache(self._h), "invalidate_norm_cache")

    def _cuda_error(self) -> str:
        return f"CUDA error {self._L.pdwt_last_cuda_error()}: {self._L.pdwt_last_cuda_error_string().decode()}"

    def set_stream(self, stream):
        """cudaStream_t (int) or torch.cuda.Stream all subsequent work is enqueued on"""
        s = getattr(stream, "cuda_stream", stream)
        _check(self._L.pdwt_wavelets_set_stream(self._h, C.c_void_p(int(s) if s else None)), "set_stream")

    def set_async(self, on: bool = True):
        """host copies of get_image/set_image/get_coeff/set_coeff are only enqueued on the object's stream (pinned host
        memory, call sync() before touching it): objects on different streams pipeline H2D / kernels / D2H"""
        _check(self._L.pdwt_wavelets_set_async(self._h, 1 if on else 0), "set_async")

    def sync(self):
        _check(self._L.pdwt_wavelets_sync(self._h), "pdwt_wavelets_sync")

    # ---- methods ------------------------------------------------------------------------------------------
    def forward(self):
        rc = self._L.pdwt_wavelets_forward(self._h)
        if rc != PDWT_ERR_STATE:
            _check(rc, "forward")

    def inverse(self):
        rc = self._L.pdwt_wavelets_inverse(self._h)
        if rc != PDWT_ERR_STATE:  # refused calls are warnings in the reference (wt.cu:274-281), not errors
            _check(rc, "inverse")

    def soft_threshold(self, beta: float, do_thresh_appcoeffs: int = 0, normalize: int = 0):
        rc = self._L.pdwt_wavelets_soft_threshold(self._h, float(beta), int(do_thresh_appcoeffs), int(normalize))
        if rc != PDWT_ERR_STATE:
            _check(rc, "soft_threshold")

    def hard_threshold(self, beta: float, do_thresh_appcoeffs: int = 0, normalize: int = 0):
        rc = self._L.pdwt_wavelets_hard_threshold(self._h, float(beta), int(do_thresh_appcoeffs), int(normalize))
        if rc != PDWT_ERR_STATE:
            _check(rc, "hard_threshold")

    def group_soft_threshold(self, beta: float, do_thresh_appcoeffs: int = 0, normalize: int = 0):
        rc = self._L.pdwt_wavelets_group_soft_threshold(self._h, float(beta), int(do_thresh_appcoeffs), int(normalize))
        if rc != PDWT_ERR_STATE:
            _check(rc, "group_soft_threshold")

    def shrink(self, beta: float, do_thresh_appcoeffs: int = 1):
        rc = self._L.pdwt_wavelets_shrink(self._h, float(beta), int(do_thresh_appcoeffs))
        if rc != PDWT_ERR_STATE:
            _check(rc, "shrink")

    def proj_linf(self, beta: float, do_thresh_appcoeffs: int = 1):
        rc = self._L.pdwt_wavelets_proj_linf(self._h, float(beta), int(do_thresh_appcoeffs))
        if rc != PDWT_ERR_STATE:
            _check(rc, "proj_linf")

    def circshift(self, sr: int, sc: int, inplace: int = 1):
        """circular shift of the image (wt.cu:365); inplace=0 leaves the result in d_tmp like the reference"""
        _check(self._L.pdwt_wavelets_circshift(self._h, int(sr), int(sc), int(inplace)), "circshift")

    def add_wavelet(self, other: "Wavelets", alpha: float = 1.0) -> int:
        """coefficients += alpha * other's (wt.cu:624-657); returns the reference's codes (0 ok, 1 / -1..-4 refused)"""
        return self._L.pdwt_wavelets_add_wavelet(self._h, other._h, float(alpha))

    @property
    def current_shift(self):
        r, c = C.c_int(), C.c_int()
        self._L.pdwt_wavelets_current_shift(self._h, C.byref(r), C.byref(c))
        return r.value, c.value

    def _norm(self, fn):
        out = np.zeros(self.batch, dtype=np.float32)
        _check(fn(self._h, out.ctypes.data_as(C.c_void_p)), "norm")
        return float(out[0]) if self.batch == 1 else out

    def norm1(self):
        return self._norm(self._L.pdwt_wavelets_norm1)

    def norm2sq(self):
        return self._norm(self._L.pdwt_wavelets_norm2sq)

    def _plane_shape(self, nr, nc):
        return (nr, nc) if self.batch == 1 else (self.batch, nr, nc)

    def get_image(self, out=None) -> np.ndarray:
        w = self.info
        if out is None:
            out = np.empty(self._plane_shape(w.Nr, w.Nc), dtype=np.float32)
        n = self._L.pdwt_wavelets_get_image(self._h, out.ctypes.data_as(C.c_void_p))
        if n != min(out.size, 0x7fffffff):   # the reference's int element count, saturated for huge batches
            raise PdwtError(f"get_image failed ({self._cuda_error()})")
        return out

    def coeff_shape(self, num: int):
        nr, nc = C.c_int(), C.c_int()
        _check(self._L.pdwt_coeff_dims(self.info, num, C.byref(nr), C.byref(nc)), "pdwt_coeff_dims")
        return nr.value, nc.value

    def get_coeff(self, num: int):
        if self.state == W_INVERSE:
            return None  # wt.cu:476-479
        out = np.empty(self._plane_shape(*self.coeff_shape(num)), dtype=np.float32)
        n = self._L.pdwt_wavelets_get_coeff(self._h, out.ctypes.data_as(C.c_void_p), num)
        if n != min(out.size, 0x7fffffff):
            raise PdwtError(f"get_coeff failed ({self._cuda_error()})")
        return out

    def set_image(self, img):
        if hasattr(img, "data_ptr"):
            _check(self._L.pdwt_wavelets_set_image(self._h, C.c_void_p(_dev_ptr(img)), 1), "set_image")
        else:
            a = np.ascontiguousarray(img, dtype=np.float32)
            _check(self._L.pdwt_wavelets_set_image(self._h, a.ctypes.data_as(C.c_void_p), 0), "set_image")

    def set_coeff(self, coeff, num: int):
        if hasattr(coeff, "data_ptr"):
            _check(self._L.pdwt_wavelets_set_coeff(self._h, C.c_void_p(_dev_ptr(coeff)), num, 1), "set_coeff")
        else:
            a = np.ascontiguousarray(coeff, dtype=np.float32)
            _check(self._L.pdwt_wavelets_set_coeff(self._h, a.ctypes.data_as(C.c_void_p), num, 0), "set_coeff")

    def set_filters_forward(self, name: str, f1, f2, f3=None, f4=None) -> int:
        """wt.cu:560-583.  Separable mode: the 1-D pair (lo, hi).  Non-separable mode: four len x len filters (the
        reference's A, H, V, D order = LL, LH, HL, HH); returns the reference's codes (0, -1 too long, -2 missing 2-D
        filters, -3 failed)."""
        f = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (f1, f2, f3, f4)]
        if f[0].ndim == 2 or f[2] is not None or not self.do_separable:
            p = [a.ctypes.data_as(_fp) if a is not None else None for a in f]
            return self._L.pdwt_wavelets_set_filters_forward_2d(self._h, name.encode(), f[0].shape[0], *p)
        return self._L.pdwt_wavelets_set_filters_forward(self._h, name.encode(), len(f[0]), f[0].ctypes.data_as(_fp),
                                                         f[1].ctypes.data_as(_fp))

    def set_filters_inverse(self, f1, f2, f3=None, f4=None) -> int:
        f = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (f1, f2, f3, f4)]
        if f[0].ndim == 2 or f[2] is not None or not self.do_separable:
            p = [a.ctypes.data_as(_fp) if a is not None else None for a in f]
            return self._L.pdwt_wavelets_set_filters_inverse_2d(self._h, *p)
        return self._L.pdwt_wavelets_set_filters_inverse(self._h, f[0].ctypes.data_as(_fp), f[1].ctypes.data_as(_fp))
