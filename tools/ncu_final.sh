# Round-end ncu captures of the kernels that dominate the C2 step (run under gpurun, one GPU).  Every capture follows a
# plain run of the same command that exited 0.  The reports are large (the module's SASS is embedded), so all but the top
# kernel's are reduced to their raw-metric CSV on the GPU box; gpurun_out/ may carry 64 MiB back.
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_ws_kernel -s 0 -c 1 -o gpurun_out/prof_r1_top_tcconv_out2 $B > gpurun_out/ncu_f1.log 2>&1
cap() {  # name, kernel regex, launches to skip
  ncu --set full --clock-control none -k regex:$2 -s $3 -c 1 -o /tmp/$1 $B > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap tcconvT2 conv_tc_ws_kernel 22
cap tcconv3x3 conv_tc_kernel 3
cap eltwise_add82 eltwise_kernel 2
cap stem conv_stem_kernel 0
cap dw3x3_first dwconv_strip_kernel 0
cap ccl_merge ccl_merge_kernel 0
cap contour_warp contour_warp_kernel 0
cap se_apply se_apply_kernel 9
ls -la gpurun_out/
