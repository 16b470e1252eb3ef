#!/usr/bin/env python3
"""Golden vectors for CUSTOM 2-D filter quadruples in non-separable mode (Wavelets::set_filters_forward/_inverse with
four filters, wt.cu:560-602; nonseparable.cu:86-106), dumped from the reference's own CUDA build on the GPU box:

    gpurun -- python tests/golden/make_golden_custom2d.py      -> gpurun_out/golden/ns2_custom_*.npz

The reference keeps ONE set of constant-memory symbols for both directions, so the only sequence that works there is
set_filters_forward -> forward -> set_filters_inverse -> inverse; that is the sequence recorded here."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import Ref, fp, load_ref  # noqa: E402

CASES = [  # name, shape, len, levels, swt
    ("ns2_custom_len6_72x96", (72, 96), 6, 2, 0),
    ("ns2_custom_len4_33x47", (33, 47), 4, 2, 0),
    ("ns2_custom_len8_64", (64, 64), 8, 1, 0),
    ("nsswt2_custom_len4_40x48", (40, 48), 4, 2, 1),
]


def quadruple(n, seed):
    """four n x n filters that are NOT outer products (a separable part plus a random perturbation)"""
    r = np.random.default_rng(seed)
    out = []
    for k in range(4):
        a, b = r.standard_normal(n), r.standard_normal(n)
        out.append((np.outer(a, b) / n + 0.05 * r.standard_normal((n, n))).astype(np.float32))
    return out


def inputs(name, shape):
    return (np.random.default_rng(abs(hash(name)) % 1000 + 1).standard_normal(shape) * 50 + 128).astype(np.float32)


def declare(L):
    L.ref_set_filters_forward.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, fp, fp, fp, fp]
    L.ref_set_filters_inverse.argtypes = [C.c_void_p, fp, fp, fp, fp]


def run_ref(L, x, n, levels, swt, F, I):
    W = Ref(L, x, "db%d" % (n // 2), levels, sep=0, swt=swt)   # any named bank of that length: only hlen is kept
    assert L.ref_set_filters_forward(W.h, b"custom", n, *[f.ctypes.data_as(fp) for f in F]) == 0
    L.ref_forward(W.h)
    out = {"meta": np.array([W.nlevels, n], np.int32)}
    for i, c in enumerate(W.coeffs()):
        out[f"c{i}"] = c
    assert L.ref_set_filters_inverse(W.h, *[f.ctypes.data_as(fp) for f in I]) == 0
    L.ref_inverse(W.h)
    out["recon"] = W.image()
    W.close()
    return out


if __name__ == "__main__":
    L = load_ref()
    declare(L)
    dst = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    for k, (name, shape, n, levels, swt) in enumerate(CASES):
        x = (np.random.default_rng(100 + k).standard_normal(shape) * 50 + 128).astype(np.float32)
        F, I = quadruple(n, 10 + k), quadruple(n, 50 + k)
        out = run_ref(L, x, n, levels, swt, F, I)
        out["x"] = x
        for j in range(4):
            out[f"F{j}"], out[f"I{j}"] = F[j], I[j]
        np.savez_compressed(os.path.join(dst, name + ".npz"), **out)
        print(name, "levels", int(out["meta"][0]), "sub-bands", sum(1 for q in out if q.startswith("c")))
