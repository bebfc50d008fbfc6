#!/bin/bash
# A/B runs of bench.py under environment knobs; one line per variant into gpurun_out/ab.txt
out=gpurun_out/ab.txt
mkdir -p gpurun_out
: > $out
run() {
  local tag="$1"; shift
  local line
  line=$(env "$@" timeout 180 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1)
  echo "$tag | $(echo "$line" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d.get("e2e",{}).get("value"), (d.get("single_stream") or {}).get("ms_per_step"))
except Exception as e: print("ERR", e)')" | tee -a $out
}
while read -r tag vars; do
  [ -z "$tag" ] && continue
  run "$tag" $vars
done < "${1:-tools/ab_variants.txt}"
