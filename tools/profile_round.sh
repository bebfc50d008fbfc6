#!/bin/bash
# ncu evidence for profiles/: launch list of one step, per-conv DRAM bytes + tensor-pipe activity, and three
# --set full captures (RPN P2 3x3 conv, layer-3 conv3 with the residual epilogue, layer-1 conv3).
# Run on the GPU box:  tools/profile_round.sh r01b
tag=${1:-r01b}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --inflight 1 --no-cpu-baseline"
export SB_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_tc|stem_|maxpool|subsample|roi_align|nms_|select_|count_eq|block_scan|compact|rank_sort|decode|write_rois|rpn_head|init_state|kpts_|box_tail|upsample2x|dense_|fill" -c 560 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_launches_$tag.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none -k regex:conv_tc_kernel -c 430 --csv --log-file $out/conv_dram_$tag.csv $B > $out/ncu_convdram_$tag.log 2>&1
i=0
for skip in 197 30 4; do
  i=$((i+1))
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel --launch-skip $skip --launch-count 1 -f \
      -o $out/prof_conv_${tag}_$i $B > $out/ncu_full_${tag}_$i.log 2>&1
done
ls -la $out/*$tag*
