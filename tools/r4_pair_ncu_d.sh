# int_bwd_d: two warps per tile vs one, at 32x32 (B = 64) and 64x64 (B = 16): ncu --set full of one launch each, raw page as csv
for SZ in "64 32" "16 64"; do set -- $SZ
  for PAIR in 1 0; do
    PTK_BWD_PAIR_ANY=1 PTK_BWD_C_PAIR=$PAIR PTK_BWD_D_PAIR=$PAIR ncu --set full --clock-control none -k regex:"int_bwd_d" -s 1 -c 1 -o /tmp/r4i_d_b$1_s$2_p$PAIR python tools/prof_step.py 4 $1 1 $2 > gpurun_out/r4i_ncu.log 2>&1
    ncu -i /tmp/r4i_d_b$1_s$2_p$PAIR.ncu-rep --page raw --csv > gpurun_out/r4i_d_b$1_s$2_p$PAIR.csv 2>/dev/null
  done
done
ls -la gpurun_out/r4i_*
