#!/bin/bash
O=gpurun_out/exp45; mkdir -p $O
PDWT_INV_TMA=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream-db3-640x600" > $O/san1.log 2>&1
grep -m3 -A6 "Illegal\|Invalid\|at pdwt" $O/san1.log | cut -c1-220 | head -40
cat > /tmp/b.py <<'P'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import pdwt_b200, oracle
x = (np.random.default_rng(0).standard_normal((3, 512, 768)) * 50 + 128).astype(np.float32)
W = pdwt_b200.Wavelets(x, "db7", 3)
W.forward(); W.inverse()
r = W.get_image()
for b in range(3):
    O = oracle.Wavelets(x[b], "db7", 3); O.forward(); O.inverse()
    d = np.abs(r[b] - O.get_image())
    bad = np.argwhere(d > 0)
    print("plane", b, "max diff", d.max(), "n bad", len(bad), "first bad", bad[:3].tolist(), "rows with bad", np.unique(bad[:,0])[:12].tolist() if len(bad) else [], "cols", np.unique(bad[:,1])[:12].tolist() if len(bad) else [])
P
PDWT_INV_TMA=1 python /tmp/b.py 2>&1 | tail -5 | cut -c1-400
