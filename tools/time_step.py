"""whole-step and per-kernel timing of forward()+inverse() for one workload; env switches (PDWT_*) are read by the
library at first use, so every variant is its own process.  usage: time_step.py [Nr Nc batch [wname levels]]"""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200
L = pdwt_b200.lib()
Nr, Nc, B = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 4096, 1)
wname = sys.argv[4] if len(sys.argv) > 4 else "db7"
levels = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rot = 4 if B * Nr * Nc <= 4096 * 4096 else 1
g = torch.Generator(device="cuda").manual_seed(0)
Ws = []
for i in range(rot):
    x = torch.randn((B, Nr, Nc) if B > 1 else (Nr, Nc), device="cuda", generator=g) * 50 + 128
    Ws.append(pdwt_b200.Wavelets(x, wname, levels))
x0 = x.clone()
for i in range(2 * rot): Ws[i % rot].forward(); Ws[i % rot].inverse()
err = float((torch.from_numpy(Ws[rot - 1].get_image()).cuda() - x0).abs().max() / x0.abs().max())
steps = 40 if rot > 1 else 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(steps): Ws[i % rot].forward(); Ws[i % rot].inverse()
e1.record(); torch.cuda.synchronize()
us = 1e3 * e0.elapsed_time(e1) / steps
L.pdwt_profile_begin()
for i in range(steps): Ws[i % rot].forward(); Ws[i % rot].inverse()
ents = (pdwt_b200.ProfileEntry * 64)()
n = L.pdwt_profile_end(ents, 64)
ks = " ".join(f"{ents[k].name.decode().replace('k_','').replace('2d','').replace('_stream','S')}={1e3 * ents[k].ms_total / ents[k].launches:.1f}" for k in range(n))
env = " ".join(f"{k[5:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("PDWT_"))
print(f"[{env or 'default'}] {B}x{Nr}x{Nc} step={us:.1f}us  {16.0*B*Nr*Nc/us/1e3:.0f} GB/s alg  err={err:.1e} | {ks}", flush=True)
