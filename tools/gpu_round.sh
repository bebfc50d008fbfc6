#!/bin/bash
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 tools/ab.sh tools/ab_variants.txt
