#!/bin/bash
# One GPU-box session of round 2: baseline (old epilogue) tests + bench + memory-kernel ncu, then the flat TMA-epilogue
# kernel: conv tests first (short timeout: a barrier bug would hang), then the model tests and the bench.
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== conv tests, flat kernel ON"
SB_TC_FLAT=1 timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 120 -x -k "conv_tc or strided" > $out/pytest_flat_conv.log 2>&1; echo "rc=$?"; tail -5 $out/pytest_flat_conv.log
echo "== full gpu suite, flat kernel OFF (baseline)"
SB_TC_FLAT=0 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s > $out/pytest_base.log 2>&1; echo "rc=$?"; tail -8 $out/pytest_base.log
echo "== bench baseline"
SB_TC_FLAT=0 timeout 400 python bench.py > $out/bench_base.json 2> $out/bench_base.err; echo "rc=$?"; tail -c 400 $out/bench_base.err
echo "== model tests, flat ON"
SB_TC_FLAT=1 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_round2.py -m gpu -q --timeout 300 -s > $out/pytest_flat_model.log 2>&1; echo "rc=$?"; tail -8 $out/pytest_flat_model.log
echo "== bench flat"
SB_TC_FLAT=1 timeout 400 python bench.py --no-cpu-baseline > $out/bench_flat.json 2> $out/bench_flat.err; echo "rc=$?"; tail -c 400 $out/bench_flat.err
SB_TC_FLAT=0 tools/profile_mem.sh r02a
nvidia-smi --query-gpu=name,memory.used --format=csv
