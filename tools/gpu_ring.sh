#!/bin/bash
for r in 0 84 44 46 28; do for th in 64 16; do echo "== RING=$r TH=$th"; PDWT_RING=$r PDWT_TH=$th python tools/time_fwd.py; done; done
echo "== default TH"; python tools/time_fwd.py
