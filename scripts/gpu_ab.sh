#!/bin/bash
# A/B of library variants on the GPU box: stage times per variant (RPSF_LIB picks the .so), then tests + bench.
# usage: scripts/gpu_ab.sh "<variant suffixes>" [tests] [bench]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
: > gpurun_out/ab.txt
for v in $1; do
  lib=regularizepsf_b200/librpsf_b200${v}.so
  [ "$v" = "default" ] && lib=regularizepsf_b200/librpsf_b200.so
  [ -f "$lib" ] || { echo "missing $lib" >> gpurun_out/ab.txt; continue; }
  if [ "$v" != "default" ]; then
    RPSF_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "config2_full_size or every_kernel_variant or float64_mode" 2>&1 | tail -1 | sed "s/^/$v parity: /" >> gpurun_out/ab.txt
  fi
  for args in ${AB_ARGS:-"8 256 2048" "1 256 2048" "32 256 2048" "8 128 1024" "1 512 8192"}; do
    RPSF_LIB=$PWD/$lib timeout 120 python scripts/stage_times.py $args 2>&1 | tail -1 >> gpurun_out/ab.txt
  done
done
cat gpurun_out/ab.txt
if [[ " $* " == *" tests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if [[ " $* " == *" bench "* ]]; then
  timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json
fi
