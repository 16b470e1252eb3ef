#!/bin/bash
for cfg in "2 4096 46" "8 2048 94"; do set -- $cfg
  B=$1 N=$2 PDWT_TM=$3 python /tmp/b.py 2>&1 | tail -1
  B=$1 N=$2 PDWT_TM=$3 PDWT_INV_TMA_EDGE=1 python /tmp/b.py 2>&1 | tail -1 | sed 's/^/EDGE /'
  B=$1 N=$2 PDWT_TM=$3 PDWT_INV_TMA_SMEM_KB=100 python /tmp/b.py 2>&1 | tail -1 | sed 's/^/SMEM100 /'
  B=$1 N=$2 PDWT_TM=$3 PDWT_INV_TMA_SMEM_KB=190 python /tmp/b.py 2>&1 | tail -1 | sed 's/^/SMEM190 /'
done
