#!/bin/bash
# ncu capture of the tensor-core convolution kernel (config 5, K = 1024); usage: profile_conv_tc.sh <tag>
TAG=${1:-convtc}
export DSPB_CONFIGS=C5
python scripts/bench_configs.py > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv_valid_tc -s 12 -c 1 -f -o gpurun_out/prof_$TAG \
    python scripts/bench_configs.py > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
