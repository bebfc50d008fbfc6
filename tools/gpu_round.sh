#!/bin/bash
# One GPU-box session of round 2 (everything that needs a B200, batched into one gpurun call).
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== dense_align + stem-fused tests"
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "dense_align" > $out/pytest_dense.log 2>&1; echo "rc=$?"; tail -3 $out/pytest_dense.log
SB_STEM_FUSED=1 timeout 400 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 150 -s -k "forward_small_fp16 or schedules_agree" > $out/pytest_stem.log 2>&1; echo "rc=$?"; tail -4 $out/pytest_stem.log
echo "== conv trace, throughput and latency schedules"
timeout 300 python tools/conv_trace.py --throughput --out $out/conv_trace_tp.json > $out/conv_trace_tp.txt 2>&1; echo "rc=$?"; tail -14 $out/conv_trace_tp.txt
timeout 300 python tools/conv_trace.py --out $out/conv_trace_lat.json > $out/conv_trace_lat.txt 2>&1; echo "rc=$?"; tail -3 $out/conv_trace_lat.txt
echo "== A/B"
timeout 900 tools/ab.sh tools/ab_variants.txt
echo "== ncu: conv DRAM bytes per step, full captures"
export SB_GRAPH=0
B="python bench.py --steps 1 --warmup 3 --inflight 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none -k regex:conv_tc -c 430 --csv --log-file $out/conv_dram_r02c.csv $B > $out/ncu_convdram_r02c.log 2>&1; echo "rc=$?"
unset SB_GRAPH
tools/profile_full.sh r02c conv-only
timeout 300 python tests/tools/dense_align_sweep.py > $out/dense_sweep.log 2>&1; echo "rc=$?"; tail -4 $out/dense_sweep.log
nvidia-smi --query-gpu=name,memory.used --format=csv
