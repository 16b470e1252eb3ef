"""level-1 inverse kernel, default (per-lane LDGSTS) against the TMA-fed variant (PDWT_INV_TMA=1), over batch sizes"""
import sys, os, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
for (B, N) in ((1, 4096), (2, 4096), (3, 4096), (4, 4096), (8, 4096), (4, 2048), (8, 2048), (16, 2048), (64, 2048), (16, 1024), (64, 1024)):
    x = torch.randn((B, N, N), device="cuda") * 50 + 128
    res = {}
    for mode in ("0", "1"):
        os.environ["PDWT_INV_TMA"] = mode
        W = pdwt_b200.Wavelets(x, "db7", 1)
        for _ in range(3):
            W.forward(); W.inverse()
        torch.cuda.synchronize()
        L.pdwt_profile_begin()
        for _ in range(10):
            W.forward(); W.inverse()
        ents = (pdwt_b200.ProfileEntry * 64)()
        n = L.pdwt_profile_end(ents, 64)
        res[mode] = {ents[k].name.decode(): round(1e3 * ents[k].ms_total / ents[k].launches, 2) for k in range(n) if 'inv' in ents[k].name.decode()}
        del W
    print(B, N, res, flush=True)
