#!/usr/bin/env python3
"""Dump golden vectors from the reference's OWN CUDA build and pin the CPU oracle against it.

Runs on the GPU box only (`gpurun -- python tests/golden/make_golden.py`): it drives
oracle/_ref/libpdwt_ref.so = the unmodified reference sources compiled for sm_100 (oracle/Makefile `ref`)
through ref_shim.cpp.  Outputs go to gpurun_out/golden/; the .npz files are then committed under
tests/golden/ together with this script.  Nothing here reads /root/reference.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import CASES, THRESH, THRESH_CASES, make_input  # noqa: E402

fp = C.POINTER(C.c_float)


def load_ref():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpdwt_ref.so"))
    L.ref_create.restype = C.c_void_p
    L.ref_create.argtypes = [fp, C.c_int, C.c_int, C.c_char_p] + [C.c_int] * 6
    for n in ("ref_destroy", "ref_forward", "ref_inverse"):
        getattr(L, n).argtypes = [C.c_void_p]
        getattr(L, n).restype = None
    for n in ("ref_soft_threshold", "ref_hard_threshold"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        getattr(L, n).restype = None
    for n in ("ref_norm1", "ref_norm2sq"):
        getattr(L, n).argtypes = [C.c_void_p]
        getattr(L, n).restype = C.c_float
    L.ref_get_image.argtypes = [C.c_void_p, fp]
    L.ref_get_coeff.argtypes = [C.c_void_p, fp, C.c_int]
    L.ref_set_image.argtypes = [C.c_void_p, fp, C.c_int]
    for n in ("ref_nlevels", "ref_hlen", "ref_ndims", "ref_state"):
        getattr(L, n).argtypes = [C.c_void_p]
    L.ref_time_fwd_inv.argtypes = [C.c_void_p, C.c_int]
    L.ref_time_fwd_inv.restype = C.c_float
    return L


class Ref:
    """thin python face of the reference class"""

    def __init__(self, L, img, wname, levels, sep=1, swt=0, ndim=2):
        self.L = L
        img = np.ascontiguousarray(img, dtype=np.float32)
        self.shape = img.shape
        self.h = L.ref_create(img.ctypes.data_as(fp), img.shape[0], img.shape[1], wname.encode(), levels, 1, sep, 0,
                              swt, ndim)
        self.nlevels, self.hlen, self.ndims = L.ref_nlevels(self.h), L.ref_hlen(self.h), L.ref_ndims(self.h)
        self.swt = swt

    def close(self):
        self.L.ref_destroy(self.h)

    def ncoeffs(self):
        return (3 if self.ndims == 2 else 1) * self.nlevels + 1

    def coeff_shape(self, num):
        nr, nc = self.shape
        if num == 0:
            scale = self.nlevels
        else:
            scale = (num - 1) // 3 + 1 if self.ndims == 2 else num
        if not self.swt:
            for _ in range(scale):
                if self.ndims == 2:
                    nr = (nr + 1) // 2
                nc = (nc + 1) // 2
        return nr, nc

    def coeff(self, num):
        out = np.empty(self.coeff_shape(num), dtype=np.float32)
        self.L.ref_get_coeff(self.h, out.ctypes.data_as(fp), num)
        return out

    def coeffs(self):
        return [self.coeff(i) for i in range(self.ncoeffs())]

    def image(self):
        out = np.empty(self.shape, dtype=np.float32)
        self.L.ref_get_image(self.h, out.ctypes.data_as(fp))
        return out

    def set_image(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        self.L.ref_set_image(self.h, img.ctypes.data_as(fp), 0)


def run_ref_case(L, x, wname, levels, sep, swt, ndim, thresh=True):
    out = {}
    W = Ref(L, x, wname, levels, sep, swt, ndim)
    out["meta"] = np.array([W.nlevels, W.hlen, W.ndims], dtype=np.int32)
    W.L.ref_forward(W.h)
    for i, c in enumerate(W.coeffs()):
        out[f"c{i}"] = c
    out["norm1"] = np.float32(W.L.ref_norm1(W.h))
    out["norm2sq"] = np.float32(W.L.ref_norm2sq(W.h))
    W.L.ref_inverse(W.h)
    out["recon"] = W.image()
    W.close()
    for tag, kind, beta, app, nrm in (THRESH if thresh else []):
        # fresh object per variant: in non-separable mode the reference's inverse() leaves the INVERSE filters in
        # the shared constant slots (wt.cu:298), so a forward() after an inverse() is wrong there.
        W = Ref(L, x, wname, levels, sep, swt, ndim)
        W.L.ref_forward(W.h)
        getattr(W.L, f"ref_{kind}_threshold")(W.h, beta, app, nrm)
        for i, c in enumerate(W.coeffs()):
            out[f"{tag}_c{i}"] = c
        out[f"{tag}_norm1"] = np.float32(W.L.ref_norm1(W.h))
        if tag == "soft":
            W.L.ref_inverse(W.h)
            out["soft_recon"] = W.image()
        W.close()
    return out


def run_oracle_case(x, wname, levels, sep, swt, ndim, thresh=True):
    import oracle
    out = {}
    W = oracle.Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
    out["meta"] = np.array([W.info.nlevels, W.info.hlen, W.info.ndims], dtype=np.int32)
    W.forward()
    for i in range(W.ncoeffs):
        out[f"c{i}"] = W.get_coeff(i)
    out["norm1"] = np.float32(W.norm1())
    out["norm2sq"] = np.float32(W.norm2sq(ref_1d_bug=1))
    W.inverse()
    out["recon"] = W.get_image()
    for tag, kind, beta, app, nrm in (THRESH if thresh else []):
        W.set_image(x)
        W.forward()
        getattr(W, f"{kind}_threshold")(beta, app, nrm)
        for i in range(W.ncoeffs):
            out[f"{tag}_c{i}"] = W.get_coeff(i)
        out[f"{tag}_norm1"] = np.float32(W.norm1())
        if tag == "soft":
            W.inverse()
            out["soft_recon"] = W.get_image()
    return out


def compare(ref, orc):
    """per-key normalised max error and bit-exactness"""
    rep = {}
    for k, r in ref.items():
        o = orc[k]
        r64, o64 = np.asarray(r, np.float64), np.asarray(o, np.float64)
        den = max(np.abs(r64).max(), 1e-30)
        rep[k] = {"err": float(np.abs(r64 - o64).max() / den),
                  "bitexact": bool(np.array_equal(np.atleast_1d(np.asarray(r)).view(np.uint8), np.atleast_1d(np.asarray(o)).view(np.uint8)))}
    return rep


def main():
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    L = load_ref()
    report = {}
    for name, shape, wname, levels, sep, swt, ndim in CASES:
        x = make_input(name, shape)
        th = name in THRESH_CASES
        ref = run_ref_case(L, x, wname, levels, sep, swt, ndim, th)
        err = L.ref_sync()
        orc = run_oracle_case(x, wname, levels, sep, swt, ndim, th)
        rep = compare(ref, orc)
        worst = max(v["err"] for v in rep.values())
        nbit = sum(v["bitexact"] for v in rep.values())
        report[name] = {"cuda_err": err, "worst_norm_err": worst, "bitexact_keys": nbit, "keys": len(rep),
                        "not_bitexact": {k: v["err"] for k, v in rep.items() if not v["bitexact"]}}
        print(f"{name:24s} cuda={err} worst={worst:.3e} bitexact {nbit}/{len(rep)}", flush=True)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **ref)

    # larger pinning cases (not stored): oracle vs reference
    big = [("big_db7_1024", (1024, 1024), "db7", 3, 1, 0, 2), ("big_sym8_swt_512", (512, 512), "sym8", 3, 1, 1, 2),
           ("big_ns_db7_512", (512, 512), "db7", 2, 0, 0, 2), ("big_haar_1023x769", (1023, 769), "haar", 4, 1, 0, 2),
           ("big_db4_1d_64x8191", (64, 8191), "db4", 5, 1, 0, 1), ("big_db7_1000x1500", (1000, 1500), "db7", 3, 1, 0, 2)]
    for name, shape, wname, levels, sep, swt, ndim in big:
        x = make_input(name, shape)
        ref = run_ref_case(L, x, wname, levels, sep, swt, ndim)
        orc = run_oracle_case(x, wname, levels, sep, swt, ndim)
        rep = compare(ref, orc)
        worst = max(v["err"] for v in rep.values())
        nbit = sum(v["bitexact"] for v in rep.values())
        report[name] = {"worst_norm_err": worst, "bitexact_keys": nbit, "keys": len(rep),
                        "not_bitexact": {k: v["err"] for k, v in rep.items() if not v["bitexact"]}}
        print(f"{name:24s} worst={worst:.3e} bitexact {nbit}/{len(rep)}", flush=True)

    # the kernel to beat: the reference's own CUDA on this B200, C2 and friends (CUDA events, 3 warm-up + 20 iters)
    timing = {}
    for name, shape, wname, levels, sep, swt in [("C2_4096_db7_L3", (4096, 4096), "db7", 3, 1, 0),
                                                 ("C3_2048_sym8_swt_L4", (2048, 2048), "sym8", 4, 1, 1),
                                                 ("C4_4096_ns_db7_L2", (4096, 4096), "db7", 2, 0, 0),
                                                 ("C5img_2048_db7_L3", (2048, 2048), "db7", 3, 1, 0),
                                                 ("haar_4096_L3", (4096, 4096), "haar", 3, 1, 0)]:
        x = make_input(name, shape)
        W = Ref(L, x, wname, levels, sep, swt, 2)
        L.ref_time_fwd_inv(W.h, 3)
        ms = L.ref_time_fwd_inv(W.h, 20) / 20
        timing[name] = {"ms_fwd_inv": ms, "Mpix_s": shape[0] * shape[1] / ms / 1e3}
        print(name, timing[name], flush=True)
        W.close()
    report["_ref_timing"] = timing
    with open(os.path.join(outdir, "pin_report.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("done in %.1fs" % (time.time() - t0))
