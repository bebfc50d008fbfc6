#!/bin/bash
# One GPU-box session of round 2 (everything that needs a B200, batched into one gpurun call).
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== full gpu suite (flat kernel on = default)"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s > $out/pytest_all.log 2>&1; echo "rc=$?"; tail -12 $out/pytest_all.log
echo "== stem fused into the GEMM producer (opt-in until verified)"
SB_STEM_FUSED=1 timeout 400 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 150 -s -k "forward_small_fp16 or schedules_agree" > $out/pytest_stem.log 2>&1; echo "rc=$?"; tail -4 $out/pytest_stem.log
echo "== bench (default)"
timeout 500 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "rc=$?"; tail -c 300 $out/bench_default.err
echo "== A/B"
timeout 1200 tools/ab.sh tools/ab_variants.txt
echo "== ncu memory kernels + full captures"
tools/profile_mem.sh r02b
tools/profile_full.sh r02b
echo "== baselines + dense_align sweep"
timeout 600 python tests/tools/baselines.py > $out/baselines.log 2>&1; echo "rc=$?"; tail -c 300 $out/baselines.log
timeout 300 python tests/tools/dense_align_sweep.py > $out/dense_sweep.log 2>&1; echo "rc=$?"; tail -4 $out/dense_sweep.log
nvidia-smi --query-gpu=name,memory.used --format=csv
