"""ncu driver: one fwd+inv of C3 (SWT sym8 L4 2048^2) and of C4 (non-separable db7 L2 4096^2)"""
import sys, torch
sys.path.insert(0, ".")
import pdwt_b200
x = torch.randn((2048, 2048), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "sym8", 4, do_swt=1)
for i in range(2):
    W.forward(); W.inverse()
y = torch.randn((4096, 4096), device="cuda") * 50 + 128
V = pdwt_b200.Wavelets(y, "db7", 2, do_separable=0)
for i in range(2):
    V.forward(); V.inverse()
torch.cuda.synchronize()
