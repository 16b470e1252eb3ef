"""PDWT_EXPERIMENTS build only: per-CTA timeline of ONE forward level kernel (globaltimer stamps, ns from the earliest
CTA start).  usage: timeline_fwd.py N   (N x N image, db7, 1 level)"""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn((N, N), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", 1)
for i in range(3):
    W.forward()
    torch.cuda.synchronize()
torch.cuda.synchronize()
buf = (C.c_ulonglong * (1024 * 8))()
L.pdwt_debug_timeline.argtypes = [C.c_void_p, C.c_int]
assert L.pdwt_debug_timeline(buf, 1024 * 8) == 0
t = np.array(buf, dtype=np.uint64).reshape(1024, 8).astype(np.int64)
t = t[t[:, 0] > 0]
n = len(t)
t0 = t[:, 0].min()
names = ["start", "synced", "prod_k0", "first_data", "pair0_done", "first_store", "cons_done", "smid"]
rel = t - t0
print(f"N={N}: {n} CTAs stamped; ns relative to the earliest CTA start")
print("stat   " + " ".join(f"{s:>11s}" for s in names))
for lab, f in (("min", np.min), ("median", np.median), ("max", np.max)):
    print(f"{lab:6s} " + " ".join(f"{int(f(rel[:, i])):11d}" for i in range(8)))
for c in (0, 1, n // 2, n - 1):
    print(f"cta{c:4d}" + " ".join(f"{int(rel[c, i]):11d}" for i in range(8)))
smid = t[:, 7]
per_sm = np.bincount(smid.astype(np.int64), minlength=148)
print("CTAs per SM: histogram", {int(k): int(v) for k, v in zip(*np.unique(per_sm, return_counts=True))})
dur = rel[:, 6] - rel[:, 1]
for k in np.unique(per_sm):
    if k == 0: continue
    sel = np.isin(smid, np.nonzero(per_sm == k)[0])
    print(f"  CTAs on SMs holding {k}: n={int(sel.sum())}  synced->done ns: min {int(dur[sel].min())} median {int(np.median(dur[sel]))} max {int(dur[sel].max())}")
if n >= 148:
    # who is slow?  by SM (both CTAs of an SM alike => an SM / memory-distance effect), by column group, by row chunk
    ncg = max(1, (N // 2 + 255) // 256)
    cg, rc = np.arange(n) % ncg, np.arange(n) // ncg
    sm = smid.astype(np.int64)
    by_sm = {k: dur[sm == k] for k in np.unique(sm)}
    two = np.array([v for v in by_sm.values() if len(v) == 2])
    if len(two) > 4:
        print(f"  the two CTAs of one SM: corr {np.corrcoef(two[:, 0], two[:, 1])[0, 1]:.2f}; SM means min {int(two.mean(1).min())} median {int(np.median(two.mean(1)))} max {int(two.mean(1).max())}")
    print("  mean by column group:", [int(dur[cg == k].mean()) for k in range(ncg)])
    q = np.quantile(rc, [0, .25, .5, .75, 1]).astype(int)
    print("  mean by row-chunk quartile:", [int(dur[(rc >= a) & (rc <= b)].mean()) for a, b in zip(q[:-1], q[1:])])
    order = np.argsort([v.mean() for v in by_sm.values()])
    keys = list(by_sm.keys())
    print("  slowest SMs:", [int(keys[i]) for i in order[-10:]], " fastest SMs:", [int(keys[i]) for i in order[:10]])
print("kernel span (first start -> last done):", int(rel[:, 6].max()), "ns;  first data after sync: median", int(np.median(rel[:, 3] - rel[:, 1])), "ns")
