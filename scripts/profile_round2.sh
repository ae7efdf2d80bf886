#!/bin/bash
# Round-2 evidence run (one GPU): sanitizer record of the two chain kernel families, ncu launch list of the bench
# command, one `ncu --set full` capture of the dominant kernel.  Every profiled command first runs plain and must exit 0.
# usage: scripts/profile_round2.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
# ---- compute-sanitizer: racecheck + synccheck, >= 4 rows per CTA so that rows overlap in the software pipeline ------
timeout 120 python scripts/run_chain.py 600 1 > $OUT/${TAG}_plain_small.log 2>&1 || exit 1
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scripts/run_chain.py 600 1 \
      > $OUT/${TAG}_sanitizer_${tool}_k_chain_spec.log 2>&1
  echo "$tool k_chain_spec: rc=$? $(grep -c 'ERROR SUMMARY' $OUT/${TAG}_sanitizer_${tool}_k_chain_spec.log) $(grep 'ERROR SUMMARY' $OUT/${TAG}_sanitizer_${tool}_k_chain_spec.log | tail -1)"
done
export DSPB_CONFIGS=C4 DSPB_C4_ROWS=8192 DSPB_C4_BLOCK=8192
timeout 120 python scripts/bench_configs.py > $OUT/${TAG}_plain_warp_small.log 2>&1 || exit 1
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scripts/bench_configs.py \
      > $OUT/${TAG}_sanitizer_${tool}_k_chain_warp.log 2>&1
  echo "$tool k_chain_warp: rc=$? $(grep 'ERROR SUMMARY' $OUT/${TAG}_sanitizer_${tool}_k_chain_warp.log | tail -1)"
done
unset DSPB_CONFIGS DSPB_C4_ROWS DSPB_C4_BLOCK
# ---- launch list of the bench command ------------------------------------------------------------------------------
timeout 300 python bench.py --steps 2 --warmup 3 --no-configs --no-e2e > $OUT/${TAG}_plain_bench.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-configs --no-e2e > $OUT/${TAG}_ncu_list.log 2>&1
# ---- full capture of the dominant kernel (one launch of 16384 rows) -----------------------------------------------------
SAVE_KERNEL=1 timeout 200 python scripts/run_chain.py 16384 3 16384 > $OUT/${TAG}_plain_spec.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_spec -s 1 -c 1 -f -o $OUT/${TAG}_prof_spec \
    python scripts/run_chain.py 16384 3 16384 > $OUT/${TAG}_ncu_spec.log 2>&1
ncu -i $OUT/${TAG}_prof_spec.ncu-rep --page raw --csv > $OUT/${TAG}_prof_spec_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_prof_spec.ncu-rep --page source --csv > $OUT/${TAG}_prof_spec_source.csv 2>/dev/null
tail -2 $OUT/${TAG}_plain_spec.log
