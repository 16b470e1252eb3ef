"""ncu driver: norm1 / soft_threshold / norm2sq on the C2 coefficient set"""
import sys, torch
sys.path.insert(0, ".")
import os
os.environ["PDWT_NORM_CACHE"] = "0"
import pdwt_b200
x = torch.randn((4096, 4096), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", 3)
W.forward()
for i in range(2):
    W.norm1(); W.soft_threshold(1.0); W.norm2sq()
torch.cuda.synchronize()
