#!/bin/bash
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== trace (stem 152)"
timeout 300 python tools/conv_trace.py --throughput --out $out/conv_trace_tp2.json > $out/conv_trace_tp2.txt 2>&1; echo "rc=$?"; sed -n 1,6p $out/conv_trace_tp2.txt | cut -c1-170; tail -3 $out/conv_trace_tp2.txt
echo "== bench default (with cpu baseline + parity) and noise A/B"
timeout 500 python bench.py > $out/bench_default2.json 2> $out/bench_default2.err; echo "rc=$?"; tail -c 300 $out/bench_default2.err; cut -c1-200 $out/bench_default2.json
timeout 900 tools/ab.sh tools/ab_variants.txt
tools/profile_mem.sh r02e
nvidia-smi --query-gpu=name,memory.used --format=csv
