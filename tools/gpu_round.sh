#!/bin/bash
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:"at_|proposal_target|rpn_loss|rcnn_loss" -c 12 --csv --log-file $out/train_kernels.csv python tests/tools/train_targets_bench.py > $out/ncu_train.log 2>&1; echo "ncu rc=$?"
grep -v "^==" $out/train_kernels.csv | grep duration | awk -F'","' '{print $5, $NF}' | cut -c1-120
