#!/bin/bash
# Round profile artefacts (run under gpurun): ncu launch list of the bench command's timed region,
# one `--set full` capture of each dominant kernel, the per-node cycle profile of the chain program.
TAG=${1:-r01}
python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_chain|dspb|cusp|zac|k_t0|k_diff" -c 400 --csv \
    --log-file gpurun_out/${TAG}_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_list_${TAG}.log 2>&1
SAVE_KERNEL=1 python scripts/run_chain.py 16384 3 16384 > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_chain_spec -s 1 -c 1 -f -o gpurun_out/${TAG}_k_chain_spec_full \
    python scripts/run_chain.py 16384 3 16384 > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}_k_chain_spec_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_chain_spec_ncu_full_raw.csv 2>/dev/null
PROFILE_PROGRAM=1 python scripts/run_chain.py 32768 2 > gpurun_out/${TAG}_per_node_cycles.txt 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
# warp-per-waveform tier (config 4)
bash scripts/profile_warp.sh ${TAG}_warp 262144
