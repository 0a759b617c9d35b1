# Round-end ncu captures of the kernels that dominate the C2 step (run under gpurun, one GPU).  Every capture follows a
# plain run of the same command that exited 0.  The reports are large (the module's SASS is embedded), so all but the top
# kernel's are reduced to their raw-metric CSV on the GPU box; gpurun_out/ may carry 64 MiB back.
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 || exit 1
# launch list of one full bench run (first 1400 launches: warm-up step + two timed steps + e2e + profile passes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1_final.csv $B > gpurun_out/ncu_list.log 2>&1
cap() {  # name, kernel regex, launches to skip, extra flags
  ncu --set full --clock-control none $4 -k regex:$2 -s $3 -c 1 -o /tmp/$1 $B > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
  if [ "$5" = keep ]; then cp /tmp/$1.ncu-rep gpurun_out/prof_r1_$1.ncu-rep; fi
  rm -f /tmp/$1.ncu-rep
}
cap tcconv3x3 conv_tc_kernel 3 "--import-source on" keep
cap tcconv_out2_packed conv_tc_ws_kernel 0
cap tcconv_out46 conv_tc_ws_kernel 21
cap tcconvT2 conv_tc_ws_kernel 22
cap stem conv_stem_kernel 0
cap eltwise_add82 eltwise_bf16x8_kernel 2
cap dw3x3_first dwconv_strip_kernel 0
cap dw5x5 dwconv_strip_kernel 6
cap gap_se gap_se_kernel 9
cap se_apply se_apply_bf16x8_kernel 9
cap concat concat_slice_vec_kernel 0
cap resize_norm resize_norm_kernel 0
cap ccl_merge ccl_merge_kernel 0
cap contour_warp contour_warp_kernel 0
ls -la gpurun_out/
