"""PDWT_EXPERIMENTS build only: per-CTA timeline of ONE cross-level forward launch (globaltimer stamps).
usage: timeline_multi.py [N batch levels]"""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
LV = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x = torch.randn((B, N, N), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", LV)
for i in range(3):
    W.forward()
    torch.cuda.synchronize()
buf = (C.c_ulonglong * (4096 * 8))()
L.pdwt_debug_timeline.argtypes = [C.c_void_p, C.c_int]
assert L.pdwt_debug_timeline(buf, 4096 * 8) == 0
t = np.array(buf, dtype=np.uint64).reshape(4096, 8).astype(np.int64)
t = t[t[:, 0] > 0]
n = len(t); t0 = t[:, 0].min()
start, synced, deps, first, done, smid = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3] - t0, t[:, 6] - t0, t[:, 7]
info = t[:, 5]; level = info & 15; ny = (info >> 4) & 0xfff; plane = (info >> 16) & 0xffff
span = done.max()
print(f"{B}x{N}x{N} L{LV}: {n} CTAs stamped, kernel span {span} ns")
dur = done - start
for l in range(LV):
    m = level == l
    if not m.any(): continue
    w = (deps - synced)[m]
    print(f"  level {l}: {int(m.sum())} items  rows {sorted(set(ny[m].tolist()))}  start {int(start[m].min())}-{int(start[m].max())}  dur med {int(np.median(dur[m]))} max {int(dur[m].max())}"
          f"  dep-wait med {int(np.median(w))} p90 {int(np.quantile(w, .9))} max {int(w.max())} sum {int(w.sum())}  start->first_data med {int(np.median((first - start)[m]))}")
bins = np.linspace(0, span, 21)
conc = [(np.minimum(done, b1) - np.maximum(start, b0)).clip(0).sum() / (b1 - b0) for b0, b1 in zip(bins[:-1], bins[1:])]
print("  resident CTAs (avg per 5% of the span):", [int(c) for c in conc])
wait = [(np.minimum(deps, b1) - np.maximum(synced, b0)).clip(0).sum() / (b1 - b0) for b0, b1 in zip(bins[:-1], bins[1:])]
print("  of which waiting for dependencies:      ", [int(c) for c in wait])
print(f"  sum of CTA durations / (span x 592) = {dur.sum() / span / 592:.2f}")
