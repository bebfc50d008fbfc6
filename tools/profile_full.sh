#!/bin/bash
# ncu --set full captures (one launch each) of the memory-/latency-bound kernels north_star names, plus the flat
# conv kernel with the TMA epilogue.  Run on the GPU box:  tools/profile_full.sh r02 ; read here with
# tools/summarize_mem.py.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --inflight 1 --no-cpu-baseline"
export SB_GRAPH=0
cap() {   # name regex skip
  timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"$2" --launch-skip $3 --launch-count 1 -f \
      -o $out/full_${tag}_$1 $B > $out/ncu_full_${tag}_$1.log 2>&1
  echo "$1 rc=$?"
}
if [ "$2" != "conv-only" ]; then
cap roialign14 roi_align_pyramid_nhwc 2
cap roialign7 roi_align_pyramid_nhwc 0
cap nms_mask nms_mask_kernel 0
cap nms_reduce nms_reduce_kernel 0
cap stem_im2col stem_im2col16 0
cap rank_sort rank_sort_kernel 0
fi
cap dense_stage0 "dense_stage_kernel<.int.0>" 0
cap class_nms class_nms_kernel 0
cap conv_flat_res "conv_tc_flat_kernel<.int.128, .int.3" 8
cap conv_flat_f16 "conv_tc_flat_kernel<.int.128, .int.6" 8
cap conv_rpn_p2 "conv_tc_kernel<.int.256, .int.4, .bool.0, .bool.0" 10
ls -la $out/full_${tag}_*.ncu-rep
