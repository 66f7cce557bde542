cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_chain|k3_stream_paired" --launch-skip 4 -c 2 -f -o gpurun_out/r02e_paired python scripts/paired_ab.py 8,256,2048 > gpurun_out/r02e_paired_ncu.log 2>&1
ncu -i gpurun_out/r02e_paired.ncu-rep --page raw --csv > gpurun_out/r02e_paired_raw.csv 2>> gpurun_out/r02e_paired_ncu.log
python scripts/ncu_raw_summary.py gpurun_out/r02e_paired_raw.csv
