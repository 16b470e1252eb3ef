"""per-kernel fwd+inv timing (library profiler) for C2"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
x = (np.random.default_rng(0).standard_normal((4096, 4096)) * 50 + 128).astype(np.float32)
Ws = [pdwt_b200.Wavelets(torch.from_numpy(x).cuda(), "db7", 3) for _ in range(4)]
for i in range(4): Ws[i].forward(); Ws[i].inverse()
torch.cuda.synchronize()
L.pdwt_profile_begin()
for i in range(20): Ws[i % 4].forward(); Ws[i % 4].inverse()
ents = (pdwt_b200.ProfileEntry * 64)()
n = L.pdwt_profile_end(ents, 64)
print({ents[k].name.decode(): round(1e3 * ents[k].ms_total / ents[k].launches, 2) for k in range(n) if 'inv' in ents[k].name.decode()})
