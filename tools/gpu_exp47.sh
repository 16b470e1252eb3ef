#!/bin/bash
cat > /tmp/b.py <<'P'
import sys, numpy as np, torch, os
sys.path.insert(0, ".")
import pdwt_b200
B = int(os.environ.get("B", "2")); N = int(os.environ.get("N", "4096"))
x = torch.randn((B, N, N), device="cuda") * 50 + 128
tot = 0; where = []
for rep in range(2):
    os.environ["PDWT_INV_TMA"] = "1"
    W = pdwt_b200.Wavelets(x, "db7", 1); W.forward(); W.inverse(); r = torch.from_numpy(W.get_image()).cuda()
    os.environ["PDWT_INV_TMA"] = "0"
    W0 = pdwt_b200.Wavelets(x, "db7", 1); W0.forward(); W0.inverse(); r0 = torch.from_numpy(W0.get_image()).cuda()
    bad = torch.nonzero((r - r0).abs() > 0)
    tot += len(bad)
    if len(bad): where.append((int(bad[0,0]), int(bad[0,1]), int(bad[0,2]), int(bad[-1,1]), int(bad[-1,2])))
    del W, W0
print(f"B={B} N={N} TM={os.environ.get('PDWT_TM')} bad={tot} first/last {where}")
P
for cfg in "1 4096 94" "1 4096 46" "2 4096 46" "2 4096 64" "2 4096 128" "2 4096 94" "2 4096 32" "4 2048 94" "2 2048 94" "8 2048 94"; do set -- $cfg; B=$1 N=$2 PDWT_TM=$3 python /tmp/b.py 2>&1 | tail -1; done
