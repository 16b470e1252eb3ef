"""the same batch as G objects on G streams (the hardware interleaves their level kernels) against ONE batched object.
usage: time_streams.py Nr Nc batch groups"""
import os, sys, torch
sys.path.insert(0, ".")
import pdwt_b200
Nr, Nc, B, G = (int(a) for a in sys.argv[1:5])
gen = torch.Generator(device="cuda").manual_seed(0)
per = B // G
streams = [torch.cuda.Stream() for _ in range(G)]
Ws = []
for g in range(G):
    x = torch.randn((per, Nr, Nc) if per > 1 else (Nr, Nc), device="cuda", generator=gen) * 50 + 128
    W = pdwt_b200.Wavelets(x, "db7", 3)
    W.set_stream(streams[g])
    Ws.append(W)
torch.cuda.synchronize()
def step(order):
    if order == "obj":       # each object's whole transform, object after object
        for W in Ws: W.forward(); W.inverse()
    else:                    # all forwards, then all inverses
        for W in Ws: W.forward()
        for W in Ws: W.inverse()
for order in ("obj", "dir"):
    for i in range(3): step(order)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for s in streams: s.wait_event(e0)
    for i in range(steps): step(order)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / steps
    env = " ".join(f"{k[5:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("PDWT_"))
    print(f"[{env or 'default'}] {B}x{Nr}x{Nc} as {G} objects on {G} streams, order={order}: step={us:.1f}us  {16.0*B*Nr*Nc/us/1e3:.0f} GB/s alg", flush=True)
