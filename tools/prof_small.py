"""ncu driver: level-1 forward of a 1024x1024 image (lone-warp regime when PDWT_TH=64)"""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
x = (np.random.default_rng(0).standard_normal((1024, 1024)) * 50 + 128).astype(np.float32)
W = pdwt_b200.Wavelets(torch.from_numpy(x).cuda(), "db7", 1)
for dbg in ("0", "1", "2"):
    os.environ["PDWT_DBG"] = dbg
    for i in range(3):
        W.forward()
torch.cuda.synchronize()
