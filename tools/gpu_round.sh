#!/bin/bash
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | tail -40
