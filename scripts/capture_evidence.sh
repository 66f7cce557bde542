#!/bin/bash
# One-GPU evidence capture for profiles/ (run through gpurun): bench line, reference arm, stage times,
# ncu launch list of the bench command, ncu --set full of one launch of each hot kernel (config 2, 8 frames), and the
# key counters of the three kernels on a single frame, config 1 and config 4.
# usage: scripts/capture_evidence.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r02x}
o=gpurun_out
mkdir -p $o
timeout 600 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_bench_n1.err
: > $o/${tag}_stage_times.txt
for args in "1 256 2048" "8 256 2048" "32 256 2048" "8 128 1024" "1 128 1024" "1 512 8192" "8 64 1024"; do
  timeout 120 python scripts/stage_times.py $args 2>&1 | tail -1 >> $o/${tag}_stage_times.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $o/${tag}_launches_bench.csv \
  python bench.py --steps 3 --warmup 3 --cpu-frames 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_stream|k2_pipelined|k3_stream" \
  --launch-skip 6 -c 3 -f -o $o/${tag}_full python scripts/profile_apply.py 8 4 > $o/${tag}_ncu.log 2>&1
ncu -i $o/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_full_raw.csv 2>> $o/${tag}_ncu.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
: > $o/${tag}_ncu_configs.txt
for cfg in "1 4 256 2048" "8 4 128 1024" "1 4 128 1024" "1 4 512 8192"; do
  echo "==== profile_apply.py $cfg (frames reps patch size)" >> $o/${tag}_ncu_configs.txt
  timeout 300 ncu --metrics $M --clock-control none -k regex:"k1_stream|k2_pipelined|k2_colfft|k3_stream" --launch-skip 6 -c 3 \
    python scripts/profile_apply.py $cfg 2>&1 | grep -E "k1_stream|k2_|k3_stream|gpu__|dram__|lts__|smsp__|sm__" | sed -E 's/\(const.*//' >> $o/${tag}_ncu_configs.txt
done
ls -la $o | tail -12
tail -c 600 $o/${tag}_bench_n1.json
cat $o/${tag}_stage_times.txt
