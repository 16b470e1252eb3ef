#!/bin/bash
# opt-in forward variant of the non-separable DWT (filter table as a kernel parameter): parity of the ns cases, C4 timing
O=gpurun_out/exp17; mkdir -p $O
export PDWT_NS_FWD_CONST=1
timeout 30 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ns2 or s0w0 or c4" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -2 $O/pytest.log
timeout 25 python - > $O/c4.txt 2>&1 <<'PY'
import torch, time, sys
sys.path.insert(0, ".")
import pdwt_b200
x = torch.randn((4096, 4096), device="cuda") * 50 + 128
Ws = [pdwt_b200.Wavelets(x, "db7", 2, do_separable=0) for _ in range(3)]
for W in Ws: W.forward(); W.inverse()
torch.cuda.synchronize()
for name, fn in (("fwd", lambda W: W.forward()), ("fwd+inv", lambda W: (W.forward(), W.inverse()))):
    t0 = time.perf_counter()
    for i in range(12): fn(Ws[i % 3])
    torch.cuda.synchronize()
    print(name, round((time.perf_counter() - t0) / 12 * 1e6, 1), "us")
PY
cat $O/c4.txt
