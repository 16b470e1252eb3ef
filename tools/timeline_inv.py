"""PDWT_EXPERIMENTS build only: per-CTA timeline of ONE level-1 inverse kernel (N x N output, db7)"""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.randn((N, N), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", 1)
for i in range(3):
    W.forward(); torch.cuda.synchronize(); W.inverse(); torch.cuda.synchronize()
buf = (C.c_ulonglong * (4096 * 8))()
L.pdwt_debug_timeline.argtypes = [C.c_void_p, C.c_int]
assert L.pdwt_debug_timeline(buf, 4096 * 8) == 0
t = np.array(buf, dtype=np.uint64).reshape(4096, 8).astype(np.int64)
t = t[(t[:, 6] > 0) & (t[:, 0] > 0)]
n = len(t)
t0 = t[:, 0].min()
start, synced, first, done, smid = t[:, 0] - t0, t[:, 1] - t0, t[:, 3] - t0, t[:, 6] - t0, t[:, 7]
print(f"N={N}: {n} CTAs; kernel span {int(done.max())} ns; start spread {int(start.max())} ns")
dur = done - synced
print("synced->done ns: min", int(dur.min()), "median", int(np.median(dur)), "max", int(dur.max()), "; first pair after sync: median", int(np.median(first - synced)))
rnd = np.arange(n) // 148
for r in range(0, int(rnd.max()) + 1, 1):
    sel = rnd == r
    print(f"  round {r:2d}: start {int(np.median(start[sel])):6d}  done median {int(np.median(done[sel])):6d}  max {int(done[sel].max()):6d}")
per_sm = np.bincount(smid.astype(np.int64), minlength=148)
print("CTAs per SM:", {int(k): int(v) for k, v in zip(*np.unique(per_sm, return_counts=True))})
