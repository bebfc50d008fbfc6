#!/bin/bash
# one GPU session: full gpu test suite, smoke, default bench (what the driver runs at round end)
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== stem fused kernel time"
SB_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"conv_tc_flat_kernel<.int.64, .int.8, .bool.0, .bool.0, .bool.1>" -c 2 --csv --log-file $out/stem_fused7.csv python bench.py --steps 1 --warmup 3 --inflight 1 --no-cpu-baseline > $out/ncu_stem.log 2>&1
grep -v "^==" $out/stem_fused7.csv | awk -F'","' '{print $(NF-2), $NF}'
echo "== bench (default)"
timeout 900 python bench.py 2>&1 | tail -1 | tee $out/bench_final.json
echo "== bench SB_STEM_FUSED=0"
SB_STEM_FUSED=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 > $out/bench_unfused.json
python -c "
import json
for f in ('bench_final','bench_unfused'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('single_pair_ms'), d.get('roofline',{}).get('frac'))
"
