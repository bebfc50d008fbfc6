#!/bin/bash
# ncu byte counts for every non-conv (memory-/latency-bound) kernel of one eager bench step:
# duration, DRAM read/write bytes, L2 bytes.  Run on the GPU box:  tools/profile_mem.sh r02a
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --inflight 1 --no-cpu-baseline"
export SB_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"stem_|maxpool|subsample|roi_align|nms_|select_|count_eq|block_scan|compact|rank_sort|decode|write_rois|rpn_head|init_state|kpts_|box_tail|upsample2x|dense_|prep_image|peer_" \
    -c 400 --csv --log-file $out/mem_$tag.csv $B > $out/ncu_mem_$tag.log 2>&1
ls -la $out/mem_$tag.csv
