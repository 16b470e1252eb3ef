"""per-kernel timing (library profiler): 1-D SWT db4 L3 on 4096 rows of 4096, forward + inverse"""
import sys, time, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
x = torch.randn((4096, 4096), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db4", 3, do_swt=1, ndim=1)
for _ in range(3):
    W.forward(); W.inverse()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    W.forward(); W.inverse()
torch.cuda.synchronize()
print(f"1-D SWT db4 L3 fwd+inv: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us")
L.pdwt_profile_begin()
for i in range(10):
    W.forward(); W.inverse()
ents = (pdwt_b200.ProfileEntry * 64)()
n = L.pdwt_profile_end(ents, 64)
print({ents[k].name.decode(): round(1e3 * ents[k].ms_total / ents[k].launches, 2) for k in range(n)})
