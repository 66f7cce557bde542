#!/bin/bash
# stage times under different environment settings: scripts/ab_env.sh "VAR=a VAR=b ..." "1 256 2048" ...
cd "$(dirname "$0")/.."
settings=$1; shift
for e in $settings; do
  for args in "$@"; do echo -n "$e  "; env $e timeout 100 python scripts/stage_times.py $args 2>&1 | tail -1; done
done
