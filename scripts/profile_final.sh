#!/bin/bash
# Final evidence run of a round (one GPU).  Every profiled command first runs plain and must exit 0.
#   1. ncu launch list of the bench command (this repository's kernels only: names start with k_)
#   2. `ncu --set full` capture of the dominant kernel k_chain_spec on the ICPC chain (one launch of 16384 rows)
#   3. the same for the C1 chain kernel and the C4 warp-tier kernel (scripts/bench_configs.py)
#   4. per-node SM cycles of the ICPC kernel (tracing build)
# usage: scripts/profile_final.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python bench.py --steps 2 --warmup 3 --no-configs --no-e2e > $OUT/${TAG}_plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-configs --no-e2e > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list: $(grep -c k_ $OUT/${TAG}_ncu_launches_bench.csv) rows"
SAVE_KERNEL=1 timeout 200 python scripts/run_chain.py 16384 3 16384 > $OUT/${TAG}_plain_spec.log 2>&1 || { echo "plain run_chain failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_spec -s 1 -c 1 -f -o $OUT/${TAG}_prof_spec \
    python scripts/run_chain.py 16384 3 16384 > $OUT/${TAG}_ncu_spec.log 2>&1
tail -1 $OUT/${TAG}_plain_spec.log
for cfg in C1 C4; do
  k=k_chain_spec; [ $cfg = C4 ] && k=k_chain_warp
  DSPB_CONFIGS=$cfg timeout 300 python scripts/bench_configs.py > $OUT/${TAG}_plain_$cfg.log 2>&1 || { echo "plain $cfg failed"; continue; }
  DSPB_CONFIGS=$cfg timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 -c 1 -f \
      -o $OUT/${TAG}_prof_$cfg python scripts/bench_configs.py > $OUT/${TAG}_ncu_$cfg.log 2>&1
  cat $OUT/${TAG}_plain_$cfg.log | cut -c1-400
done
PROFILE_PROGRAM=1 timeout 600 python scripts/run_chain.py 16384 3 16384 > $OUT/${TAG}_per_node_cycles.txt 2>&1
ls -la $OUT/${TAG}_prof_*.ncu-rep
