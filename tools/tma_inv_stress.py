"""stress check of the TMA-fed inverse level kernel (PDWT_INV_TMA=1) against the default one: bit-exact or not,
over repetitions, batch sizes and chunk heights.  usage: B=2 N=4096 REPS=6 [PDWT_TM=46] python tools/tma_inv_stress.py"""
import sys, numpy as np, torch, os
sys.path.insert(0, ".")
import pdwt_b200
B = int(os.environ.get("B", "2")); N = int(os.environ.get("N", "4096")); REPS = int(os.environ.get("REPS", "4"))
wname = os.environ.get("WNAME", "db7"); L = int(os.environ.get("LEVELS", "1"))
x = torch.randn((B, N, N), device="cuda") * 50 + 128
tot = 0; where = []
for rep in range(REPS):
    os.environ["PDWT_INV_TMA"] = "1"
    W = pdwt_b200.Wavelets(x, wname, L); W.forward(); W.inverse(); r = torch.from_numpy(W.get_image()).cuda()
    os.environ["PDWT_INV_TMA"] = "0"
    W0 = pdwt_b200.Wavelets(x, wname, L); W0.forward(); W0.inverse(); r0 = torch.from_numpy(W0.get_image()).cuda()
    bad = torch.nonzero((r - r0).abs() > 0)
    tot += len(bad)
    if len(bad): where.append((int(bad[0,0]), int(bad[0,1]), int(bad[0,2]), int(bad[-1,1]), int(bad[-1,2])))
    del W, W0
print(f"B={B} N={N} {wname} L{L} TM={os.environ.get('PDWT_TM')} reps={REPS} bad={tot} first/last {where}")
