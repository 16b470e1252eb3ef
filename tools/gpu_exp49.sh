#!/bin/bash
O=gpurun_out/exp49; mkdir -p $O
cat > /tmp/c.py <<'P'
import sys, torch, os
sys.path.insert(0, ".")
import pdwt_b200
x = torch.randn((8, 2048, 2048), device="cuda") * 50 + 128
W = pdwt_b200.Wavelets(x, "db7", 1); W.forward(); torch.cuda.synchronize(); W.inverse(); torch.cuda.synchronize()
print("done")
P
PDWT_INV_TMA=1 PDWT_TM=94 timeout 500 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/c.py > $O/race.log 2>&1
grep -c "hazard" $O/race.log; grep -m4 -B2 -A12 "hazard" $O/race.log | cut -c1-200 | head -70; tail -3 $O/race.log
