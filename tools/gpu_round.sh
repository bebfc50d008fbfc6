#!/bin/bash
# One GPU-box session of round 2 (everything that needs a B200, batched into one gpurun call).
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== tests touched by the last changes"
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --timeout 300 -k "roi_align or forward_small_fp16 or stem" > $out/pytest_quick.log 2>&1; echo "rc=$?"; tail -3 $out/pytest_quick.log
echo "== A/B"
timeout 900 tools/ab.sh tools/ab_variants.txt
echo "== ncu full captures (conv flat, dense)"
tools/profile_full.sh r02d conv-only
nvidia-smi --query-gpu=name,memory.used --format=csv
