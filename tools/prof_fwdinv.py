"""tiny driver for ncu: a few forward()+inverse() of the C2 image (4096^2 db7 L3)"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = (np.random.default_rng(0).standard_normal((4096, 4096)) * 50 + 128).astype(np.float32)
Ws = [pdwt_b200.Wavelets(torch.from_numpy(x).cuda(), "db7", 3) for _ in range(2)]
for i in range(n):
    W = Ws[i % 2]
    W.forward(); W.inverse()
torch.cuda.synchronize()
