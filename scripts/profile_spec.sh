#!/bin/bash
# ncu capture of the specialised ICPC chain kernel (one launch of 16384 rows); usage: profile_spec.sh <tag>
TAG=${1:-spec}
SAVE_KERNEL=1 python scripts/run_chain.py 16384 3 16384 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_chain_spec -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python scripts/run_chain.py 16384 3 16384 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
