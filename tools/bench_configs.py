"""BASELINE.json configs[2] (C3: 2-D SWT sym8 L4 2048^2) and configs[3] (C4: non-separable db7 L2 4096^2 with
soft_threshold + norm1) -- ours against the reference's own CUDA build (oracle/_ref/libpdwt_ref.so) on the same GPU.
Device-resident data, wall clock around a synchronised loop (norm1 returns a host scalar in both, so every C4 iteration
synchronises anyway).  Prints one JSON line per config."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200

fp = C.POINTER(C.c_float)
R = None
so = os.path.join("oracle", "_ref", "libpdwt_ref.so")
if os.path.exists(so):
    R = C.CDLL(so)
    R.ref_create.restype = C.c_void_p
    R.ref_create.argtypes = [fp, C.c_int, C.c_int, C.c_char_p] + [C.c_int] * 6
    for n in ("ref_forward", "ref_inverse", "ref_destroy"):
        getattr(R, n).argtypes = [C.c_void_p]
    R.ref_soft_threshold.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
    R.ref_norm1.argtypes = [C.c_void_p]
    R.ref_norm1.restype = C.c_float
    R.ref_set_image.argtypes = [C.c_void_p, fp, C.c_int]


def rnd(shape, seed):
    return (np.random.default_rng(seed).standard_normal(shape) * 50 + 128).astype(np.float32)


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


def run(tag, shape, wname, levels, sep, swt, seq, iters, bytes_per_px, ndim=2):
    x = rnd(shape, 0)
    out = {"config": tag, "shape": list(shape), "wavelet": wname, "levels": levels, "iters": iters}
    W = pdwt_b200.Wavelets(torch.from_numpy(x).cuda(), wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
    L = pdwt_b200.lib()
    l0 = L.pdwt_launch_count()
    us = timeit(lambda: seq(W, None), iters)
    out["ours_us"] = round(us, 1)
    out["ours_launches_per_iter"] = (L.pdwt_launch_count() - l0) // (iters + 3)
    out["ours_alg_GBs"] = round(bytes_per_px * x.size / us / 1e3, 1)
    if R is not None:
        h = R.ref_create(x.ctypes.data_as(fp), shape[0], shape[1], wname.encode(), levels, 1, sep, 0, swt, ndim)
        us_r = timeit(lambda: seq(None, h), iters)
        out["ref_us"] = round(us_r, 1)
        out["speedup"] = round(us_r / us, 2)
        R.ref_destroy(h)
    print(json.dumps(out), flush=True)


def seq_fwd_inv(W, h):
    if W is not None:
        W.forward(); W.inverse()
    else:
        R.ref_forward(h); R.ref_inverse(h)


def seq_c4(W, h):   # README.md:90-103 -- forward, norm1, soft_threshold, norm1, inverse
    if W is not None:
        W.forward(); W.norm1(); W.soft_threshold(10.0, 0, 0); W.norm1(); W.inverse()
    else:
        R.ref_forward(h); R.ref_norm1(h); R.ref_soft_threshold(h, 10.0, 0, 0); R.ref_norm1(h); R.ref_inverse(h)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4", "c4db2", "haar", "c2seq", "dwt1d", "swt1d", "haar1d", "nsswt", "odd"]
    if "c3" in which:
        run("C3 swt sym8 L4 2048^2 fwd+inv", (2048, 2048), "sym8", 4, 1, 1, seq_fwd_inv, 10, 112)
    if "c4" in which:
        run("C4 nonseparable db7 L2 4096^2 fwd,norm1,soft,norm1,inv", (4096, 4096), "db7", 2, 0, 0, seq_c4, 5, 31.5)
    if "c4db2" in which:
        run("C4 nonseparable db2 L2 4096^2 fwd,norm1,soft,norm1,inv", (4096, 4096), "db2", 2, 0, 0, seq_c4, 10, 31.5)
    if "haar" in which:
        run("haar L3 4096^2 fwd+inv", (4096, 4096), "haar", 3, 1, 0, seq_fwd_inv, 20, 16)
    if "c2seq" in which:
        run("separable db7 L3 4096^2 fwd,norm1,soft,norm1,inv", (4096, 4096), "db7", 3, 1, 0, seq_c4, 20, 31.5)
    if "dwt1d" in which:
        run("1-D DWT db7 L3, 4096 rows of 4096 fwd+inv", (4096, 4096), "db7", 3, 1, 0, seq_fwd_inv, 20, 16, ndim=1)
    if "swt1d" in which:
        run("1-D SWT db4 L3, 4096 rows of 4096 fwd+inv", (4096, 4096), "db4", 3, 1, 1, seq_fwd_inv, 10, 64, ndim=1)
    if "haar1d" in which:
        run("1-D haar L3, 4096 rows of 4096 fwd+inv", (4096, 4096), "haar", 3, 1, 0, seq_fwd_inv, 20, 16, ndim=1)
    if "nsswt" in which:
        run("non-separable SWT db2 L2 1024^2 fwd+inv", (1024, 1024), "db2", 2, 0, 1, seq_fwd_inv, 10, 64)
    if "odd" in which:
        run("separable db7 L3 4095x4097 (odd sizes) fwd+inv", (4095, 4097), "db7", 3, 1, 0, seq_fwd_inv, 20, 16)
