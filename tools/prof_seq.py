"""per-kernel timing (library profiler) of the iterative-reconstruction sequence on C2:
forward, norm1, soft_threshold, norm1, inverse (README.md:90-103)"""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
x = torch.randn((4096, 4096), device="cuda") * 50 + 128
Ws = [pdwt_b200.Wavelets(x, "db7", 3) for _ in range(4)]
def seq(W):
    W.forward(); a = W.norm1(); W.soft_threshold(10.0); b = W.norm1(); W.inverse()
for W in Ws: seq(W)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20): seq(Ws[i % 4])
torch.cuda.synchronize()
print(f"sequence: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us")
for name, fn in (("norm1", lambda W: W.norm1()), ("soft", lambda W: W.soft_threshold(10.0)), ("norm2sq", lambda W: W.norm2sq())):
    for W in Ws: W.forward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20): fn(Ws[i % 4])
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us (wall, incl. sync)")
L.pdwt_profile_begin()
for i in range(20): seq(Ws[i % 4])
ents = (pdwt_b200.ProfileEntry * 64)()
n = L.pdwt_profile_end(ents, 64)
print({ents[k].name.decode(): round(1e3 * ents[k].ms_total / ents[k].launches, 2) for k in range(n)})
