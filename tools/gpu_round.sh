#!/bin/bash
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== CG2 direct: conv tests"
SB_TC_CG2=1 SB_CG2_DIRECT=1 timeout 240 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 90 -x -k "conv_tc_fp16" > $out/pytest_cg2d.log 2>&1; echo "rc=$?"; tail -4 $out/pytest_cg2d.log
for v in "SB_X=0" "SB_TC_CG2=1" "SB_TC_CG2=1 SB_CG2_DIRECT=1"; do
  echo "== trace $v"
  env $v timeout 200 python tools/conv_trace.py --throughput --out $out/conv_trace_x.json > $out/conv_trace_x.txt 2>&1; echo "rc=$?"
  grep -E "^ *(38|111|123) " $out/conv_trace_x.txt | cut -c1-100; grep -E "^3x3|all convs" $out/conv_trace_x.txt | cut -c1-60
done
