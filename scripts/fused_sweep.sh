#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r02_fused_solo.txt
: > $out
run() { echo "== $*" >> $out; env "$@" FUSED_STATS=1 timeout 100 python scripts/fused_check.py ${CASES:-8,256,2048,2048} 2>&1 | tail -3 >> $out; }
run RPSF_FUSED_CHUNKS=8 RPSF_FUSED_SOLO=1
run RPSF_FUSED_CHUNKS=8 RPSF_FUSED_SOLO=2
run RPSF_FUSED_CHUNKS=8 RPSF_FUSED_SOLO=21
run RPSF_FUSED_CHUNKS=8 RPSF_FUSED_SOLO=22
CASES="8,256,2048,2048 1,256,2048,2048" run RPSF_FUSED_CHUNKS=8 N1=46 N2=62 N3=40
CASES="8,256,2048,2048 1,256,2048,2048" run RPSF_FUSED_CHUNKS=8 RPSF_FUSED_SPLIT=40,68,40 N1=40 N2=68 N3=40
cat $out
