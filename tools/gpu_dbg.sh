#!/bin/bash
for dbg in 0 1 2 16; do for th in 64 16; do
  echo "== PDWT_DBG=$dbg TH=$th"
  PDWT_DBG=$dbg PDWT_TH=$th timeout 300 python - <<'P'
import sys; sys.path.insert(0,'.')
import json,subprocess
import bench
P
  PDWT_DBG=$dbg PDWT_TH=$th timeout 300 python tools/time_fwd.py
done; done
