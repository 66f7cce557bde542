#!/bin/bash
# stage times of library variants at explicit (B P HW) triples: scripts/ab_small.sh "<variants>" "1 256 2048" "2 256 2048" ...
cd "$(dirname "$0")/.."
variants=$1; shift
for v in $variants; do
  lib=regularizepsf_b200/librpsf_b200${v}.so; [ "$v" = "default" ] && lib=regularizepsf_b200/librpsf_b200.so
  for args in "$@"; do RPSF_LIB=$PWD/$lib timeout 100 python scripts/stage_times.py $args 2>&1 | tail -1; done
done
