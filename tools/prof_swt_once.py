"""ncu driver: C3 (SWT sym8 L4 2048^2), two forward + inverse passes"""
import sys, torch
sys.path.insert(0, ".")
import pdwt_b200
x = torch.randn((2048, 2048), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "sym8", 4, do_swt=1)
for i in range(2):
    W.forward(); W.inverse()
torch.cuda.synchronize()
