#!/bin/bash
# repeat a short bench under each environment setting and report failures:  tools/gpu_flaky.sh N "A=1 B=0" ...
n=$1; shift
for kv in "$@"; do
  ok=0; bad=0
  for i in $(seq $n); do
    if env $kv timeout 120 python bench.py --steps 300 --no-cpu-baseline --no-e2e > /tmp/fl.json 2> /tmp/fl.err; then ok=$((ok+1)); else bad=$((bad+1)); tail -2 /tmp/fl.err | cut -c1-200; fi
  done
  echo "$kv: ok=$ok bad=$bad"
done
