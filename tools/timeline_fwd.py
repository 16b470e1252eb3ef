"""PDWT_EXPERIMENTS build only: per-CTA timeline of ONE forward level kernel (globaltimer stamps, ns from the earliest
CTA start).  usage: timeline_fwd.py N   (N x N image, db7, 1 level)"""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn((N, N), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", 1)
for i in range(3): W.forward()
torch.cuda.synchronize()
buf = (C.c_ulonglong * (1024 * 8))()
L.pdwt_debug_timeline.argtypes = [C.c_void_p, C.c_int]
assert L.pdwt_debug_timeline(buf, 1024 * 8) == 0
t = np.array(buf, dtype=np.uint64).reshape(1024, 8).astype(np.int64)
t = t[t[:, 0] > 0]
n = len(t)
t0 = t[:, 0].min()
names = ["start", "synced", "prod_k0", "first_data", "pair0_done", "first_store", "cons_done", "prod_done"]
rel = t - t0
print(f"N={N}: {n} CTAs stamped; ns relative to the earliest CTA start")
print("stat   " + " ".join(f"{s:>11s}" for s in names))
for lab, f in (("min", np.min), ("median", np.median), ("max", np.max)):
    print(f"{lab:6s} " + " ".join(f"{int(f(rel[:, i])):11d}" for i in range(8)))
for c in (0, 1, n // 2, n - 1):
    print(f"cta{c:4d}" + " ".join(f"{int(rel[c, i]):11d}" for i in range(8)))
