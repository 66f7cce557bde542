#!/bin/bash
# round-2 probe: packed fp32 pipe rates, and stage times where the spectrum workspace is L2-resident
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
[ -x scripts/micro/ffma2_bench ] || nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/ffma2_bench scripts/micro/ffma2_bench.cu
scripts/micro/ffma2_bench > gpurun_out/r02_ffma2.txt 2>&1
cat gpurun_out/r02_ffma2.txt
: > gpurun_out/r02_l2_probe.txt
for args in "8 256 2048" "1 256 2048" "1 256 1024" "2 256 1024" "4 256 1024" "1 256 1536" "8 128 1024" "2 128 1024" "1 128 1024" "1 512 8192" "1 512 2048"; do
  timeout 120 python scripts/stage_times.py $args 2>&1 | tail -1 >> gpurun_out/r02_l2_probe.txt
done
cat gpurun_out/r02_l2_probe.txt
