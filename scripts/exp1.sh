for X in "" "-DDSPB_X_SKIP_P1" "-DDSPB_X_SKIP_P2" "-DDSPB_X_SKIP_P3" "-DDSPB_X_SKIP_P1 -DDSPB_X_SKIP_P2 -DDSPB_X_SKIP_P3"; do
  echo "== $X"; DSPEED_B200_NVCC_EXTRA="$X" PROFILE_PROGRAM=1 python scripts/run_chain.py 32768 3 2>&1 | grep -E "pass 2|conv_seg_group"
done
