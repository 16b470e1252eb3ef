#!/usr/bin/env python3
"""Golden vectors of the proximal operators from the reference's OWN CUDA build (oracle/_ref/libpdwt_ref.so), and the
pinning of the CPU oracle against them.  GPU box only: `gpurun -- python tests/golden/make_golden_prox.py`; outputs in
gpurun_out/golden_prox/, then committed under tests/golden/.  Nothing here reads /root/reference."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import Ref, load_ref, fp  # noqa: E402
from prox_cases import ADD_ALPHA, PROX_CASES, PROX_OPS, SHIFTS, prox_input  # noqa: E402


def main():
    import oracle
    pin_only = "--pin-only" in sys.argv   # CPU: re-pin the oracle against the committed vectors, no reference run
    if pin_only:
        return pin(oracle, os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests", "golden"))
    L = load_ref()
    L.ref_group_soft_threshold.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
    L.ref_shrink.argtypes = [C.c_void_p, C.c_float, C.c_int]
    L.ref_proj_linf.argtypes = [C.c_void_p, C.c_float, C.c_int]
    L.ref_circshift.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ref_add_wavelet.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
    out_dir = os.path.join(ROOT, "gpurun_out", "golden_prox")
    os.makedirs(out_dir, exist_ok=True)
    report = {}
    for name, shape, wname, levels, sep, swt, ndim in PROX_CASES:
        x, y = prox_input(shape, 7), prox_input(shape, 8)
        g = {}
        for tag, meth, args in PROX_OPS:
            W = Ref(L, x, wname, levels, sep, swt, ndim)
            L.ref_forward(W.h)
            getattr(L, "ref_" + meth)(W.h, *args)
            for i, c in enumerate(W.coeffs()):
                g[f"{tag}_c{i}"] = c
            W.close()
        W, W2 = Ref(L, x, wname, levels, sep, swt, ndim), Ref(L, y, wname, levels, sep, swt, ndim)
        L.ref_forward(W.h); L.ref_forward(W2.h)
        g["add_rc"] = np.int32(L.ref_add_wavelet(W.h, W2.h, ADD_ALPHA))
        for i, c in enumerate(W.coeffs()):
            g[f"add_c{i}"] = c
        W.close(); W2.close()
        for k, (sr, sc) in enumerate(SHIFTS):
            W = Ref(L, x, wname, levels, sep, swt, ndim)
            L.ref_circshift(W.h, sr, sc, 1)
            g[f"shift{k}"] = W.image()
            W.close()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **g)
    pin(oracle, out_dir, out_dir)


def ref_1d_add_touched(shape, levels, k, n):
    """w_add_coeffs_1d (common.cu:515-526) halves the width with FLOOR, so on odd-sized levels the reference's axpy stops
    short of the sub-band's end: number of leading elements of sub-band k it updates"""
    Nr, Nc = shape
    for _ in range(levels if k == 0 else k):
        Nc //= 2
    return min(n, Nr * Nc)


def pin(oracle, vec_dir, out_dir):
    """which contraction variant of the group threshold matches the reference, everything else bit-exact"""
    report = {}
    for name, shape, wname, levels, sep, swt, ndim in PROX_CASES:
        x, y = prox_input(shape, 7), prox_input(shape, 8)
        g = np.load(os.path.join(vec_dir, name + ".npz"))
        bad, variant_ok = {}, {}
        for tag, meth, args in PROX_OPS:
            variants = (0, 1, 2) if meth == "group_soft_threshold" else (None,)
            for v in variants:
                O = oracle.Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
                O.forward()
                getattr(O, meth)(*args, **({"variant": v} if v is not None else {}))
                ok = all(np.array_equal(O.get_coeff(i).view(np.uint32), g[f"{tag}_c{i}"].view(np.uint32))
                         for i in range(O.ncoeffs))
                if v is None:
                    if not ok:
                        bad[tag] = "mismatch"
                else:
                    variant_ok.setdefault(tag, {})[v] = bool(ok)
        O, O2 = (oracle.Wavelets(a, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim) for a in (x, y))
        O.forward(); O2.forward()
        rc = O.add_wavelet(O2, ADD_ALPHA)
        for i in range(O.ncoeffs):
            a, b = O.get_coeff(i).ravel(), g[f"add_c{i}"].ravel()
            n = ref_1d_add_touched(shape, O.info.nlevels, i, a.size) if (ndim == 1 and not swt) else a.size
            if rc != int(g["add_rc"]) or not np.array_equal(a[:n].view(np.uint32), b[:n].view(np.uint32)):
                bad["add"] = f"rc {rc} vs {int(g['add_rc'])}, sub-band {i}, max abs diff {float(np.abs(a[:n] - b[:n]).max()):.3e}"
        for k, (sr, sc) in enumerate(SHIFTS):
            O = oracle.Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
            O.circshift(sr, sc, 1)
            if not np.array_equal(O.get_image().view(np.uint32), g[f"shift{k}"].view(np.uint32)):
                bad[f"shift{k}"] = "mismatch"
        report[name] = {"not_bitexact": bad, "group_variants": variant_ok}
        print(name, report[name], flush=True)
    report["_note"] = ("1-D add_wavelet: compared on the elements the reference updates (its 1-D variant halves with floor, "
                       "common.cu:518, and skips the tail of odd-sized levels; oracle and CUDA path update the whole sub-band)")
    json.dump(report, open(os.path.join(out_dir, "prox_pin_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
