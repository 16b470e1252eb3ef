"""ncu driver: forward()+inverse() of a batch of 16 images of 2048^2 (throughput regime: ~1000 CTAs per SM-wave)"""
import sys, torch
sys.path.insert(0, ".")
import pdwt_b200
x = torch.randn((16, 2048, 2048), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", 3)
for i in range(2):
    W.forward(); W.inverse()
torch.cuda.synchronize()
