#!/bin/bash
# one GPU session: full gpu test suite, smoke, default bench (what the driver runs at round end) + train-stage timing
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== train-stage kernels"
timeout 600 python tests/tools/train_targets_bench.py 2>&1 | tail -12
echo "== bench (default)"
timeout 900 python bench.py 2>&1 | tail -1 | tee $out/bench_final.json | cut -c1-300
