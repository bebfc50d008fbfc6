#!/bin/bash
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== full gpu suite"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $out/pytest_all2.log 2>&1; echo "rc=$?"; tail -6 $out/pytest_all2.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default)"
timeout 500 python bench.py > $out/bench_default3.json 2> $out/bench_default3.err; echo "rc=$?"; tail -c 300 $out/bench_default3.err; cut -c1-120 $out/bench_default3.json
