#!/bin/bash
# A/B of two builds of the library on ONE box: tools/ab.sh <libA.so> <libB.so>  (bench value + per-launch event table)
mkdir -p gpurun_out
for tag in A B A B; do
  lib=$1; [ $tag = B ] && lib=$2
  PLNR_LIB=$lib python bench.py --steps 300 --warmup 5 --no-cpu-baseline --dump gpurun_out/ab_$tag.json > gpurun_out/ab_$tag.line 2>/dev/null
  python - <<PY
import json
d = json.load(open('gpurun_out/ab_$tag.line')); t = json.load(open('gpurun_out/ab_$tag.json'))['table']
print('$tag', '$lib', round(d['value']), 'img/s', round(d['ms_per_step'] * 1e3, 1), 'us/step |', ' '.join('%.0f' % (r['ms'] * 1e3) for r in t), flush=True)
PY
done
