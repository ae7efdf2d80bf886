#!/bin/bash
# ncu capture of the warp-per-waveform SiPM chain kernel (config 4); usage: profile_warp.sh <tag> [rows_per_launch]
TAG=${1:-warp}
BLK=${2:-262144}
export DSPB_CONFIGS=C4 DSPB_C4_ROWS=$((BLK*2)) DSPB_C4_BLOCK=$BLK
python scripts/bench_configs.py > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_chain_warp -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python scripts/bench_configs.py > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
