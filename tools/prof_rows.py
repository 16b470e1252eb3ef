"""ncu driver: 1-D DWT db7 L3 on 4096 rows of 4096"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
x = torch.randn((4096, 4096), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", 3, ndim=1)
for i in range(3):
    W.forward(); W.inverse()
torch.cuda.synchronize()
