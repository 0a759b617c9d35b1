# Round-2 ncu evidence for the C2 step (run under gpurun, one GPU).  Every capture follows a plain run of the same
# command that exited 0.  Reports are reduced to their raw-metric CSV on the GPU box (the module's SASS is embedded in
# every .ncu-rep); the summary is made here with tools/summarize_ncu_raw.py / tools/summarize_launches.py.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r2_plain_bench.log 2>&1 || exit 1
# launch list: warm-up (4 steps) + 2 timed steps of the streaming path + e2e + sync + profile passes
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_final.csv $B > gpurun_out/r2_ncu_list.log 2>&1
ONLY="$1"   # optional: space-separated capture names to (re)do, e.g. "pre_stem concat"
cap() {  # name, kernel regex, launches to skip
  if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $1 "; then return; fi
  ncu --set full --clock-control none -k regex:$2 -s $3 -c 1 -o /tmp/$1 $B > gpurun_out/r2_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_$1.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap pre_stem pre_stem_mma_kernel 0
cap ws_out2_packed conv_tc_ws_kernel 0
cap ws_out46_sums conv_tc_ws_kernel 17
cap ws_halo3x3_out58 conv_tc_ws_kernel 21
cap ws_convT2_tail conv_tc_ws_kernel 23
cap dw3x3_first dwconv_strip_kernel 0
cap dw5x5 dwconv_strip_kernel 6
cap concat concat_multi_bf16_kernel 0
cap fpn_add82 fpn_add_bf16_kernel 2
cap se_fc_sums se_fc_sums_kernel 9
cap ccl_merge ccl_merge4_kernel 0
cap contour_warp contour_warp_kernel 0
cap score score_kernel 0
ls -la gpurun_out/ | tail -30
