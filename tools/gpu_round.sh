#!/bin/bash
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== CG2 conv tests (short timeout: a protocol bug would hang)"
SB_TC_CG2=1 timeout 240 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 90 -x -k "conv_tc_fp16" > $out/pytest_cg2.log 2>&1; echo "rc=$?"; tail -6 $out/pytest_cg2.log
echo "== CG2 model tests"
SB_TC_CG2=1 timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 120 -x -k "forward_small_fp16 or schedules_agree" > $out/pytest_cg2b.log 2>&1; echo "rc=$?"; tail -4 $out/pytest_cg2b.log
echo "== A/B"
timeout 600 tools/ab.sh tools/ab_variants.txt
nvidia-smi --query-gpu=name,memory.used --format=csv
