"""ncu driver: batched 1-D SWT db4 L3 on 4096 rows of 4096, two forward + inverse passes"""
import sys, torch
sys.path.insert(0, ".")
import pdwt_b200
x = torch.randn((4096, 4096), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db4", 3, do_swt=1, ndim=1)
for i in range(2):
    W.forward(); W.inverse()
torch.cuda.synchronize()
