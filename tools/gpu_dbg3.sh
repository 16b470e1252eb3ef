#!/bin/bash
for th in 64 16; do echo "== TH=$th"; PDWT_DBG=3 PDWT_TH=$th python tools/time_fwd.py 2>&1 | grep -E "DBG3|k_fwd" | sort | uniq -c | sort -rn | head -8; done
