"""PDWT_EXPERIMENTS build only: per-CTA timeline of ONE batched forward level-1 kernel (globaltimer stamps).
usage: timeline_batch.py [N batch]"""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
x = torch.randn((B, N, N), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", 1)
for i in range(3):
    W.forward()
    torch.cuda.synchronize()
buf = (C.c_ulonglong * (4096 * 8))()
L.pdwt_debug_timeline.argtypes = [C.c_void_p, C.c_int]
assert L.pdwt_debug_timeline(buf, 4096 * 8) == 0
t = np.array(buf, dtype=np.uint64).reshape(4096, 8).astype(np.int64)
t = t[t[:, 0] > 0]
n = len(t); t0 = t[:, 0].min()
start, synced, first, done, smid = t[:, 0] - t0, t[:, 1] - t0, t[:, 3] - t0, t[:, 6] - t0, t[:, 7]
span = done.max()
print(f"{B}x{N}x{N}: {n} CTAs stamped, kernel span {span} ns")
dur = done - start
order = np.argsort(start)
for lab, sel in (("first 592", order[:592]), ("middle", order[592:n - 592]), ("last 592", order[n - 592:])):
    if len(sel) == 0: continue
    print(f"  {lab:10s}: start {int(start[sel].min())}-{int(start[sel].max())}  dur min/med/max {int(dur[sel].min())}/{int(np.median(dur[sel]))}/{int(dur[sel].max())}"
          f"  start->first_data med {int(np.median(first[sel] - start[sel]))}  done {int(done[sel].min())}-{int(done[sel].max())}")
# concurrency over time
bins = np.linspace(0, span, 21)
conc = [(np.minimum(done, b1) - np.maximum(start, b0)).clip(0).sum() / (b1 - b0) for b0, b1 in zip(bins[:-1], bins[1:])]
print("  resident CTAs (avg per 5% of the span):", [int(c) for c in conc])
per_sm_busy = np.zeros(148)
for s in range(148):
    m = smid == s
    per_sm_busy[s] = dur[m].sum()
print(f"  CTA-time per SM (sum of durations / span): min {per_sm_busy.min()/span:.2f} med {np.median(per_sm_busy)/span:.2f} max {per_sm_busy.max()/span:.2f}")
print(f"  CTAs per SM: min {np.bincount(smid, minlength=148).min()} max {np.bincount(smid, minlength=148).max()}")
