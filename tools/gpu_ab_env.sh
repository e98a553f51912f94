#!/bin/bash
# A/B of environment switches on ONE box:  tools/gpu_ab_env.sh "VAR=a" "VAR=b" ...   (each run: bench value + per-launch table)
mkdir -p gpurun_out
i=0
for rep in 1 2; do
for kv in "$@"; do
  i=$((i+1))
  env $kv timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-e2e --dump gpurun_out/ab_env_$i.json > gpurun_out/ab_env_$i.line 2> gpurun_out/ab_env_$i.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_env_$i.line')); t = json.load(open('gpurun_out/ab_env_$i.json'))['table']
    print('$kv', round(d['value']), 'img/s', round(d['ms_per_step'] * 1e3, 1), 'us/step |', ' '.join('%.0f' % (r['ms'] * 1e3) for r in t), flush=True)
except Exception as e:
    print('$kv failed', e); print(open('gpurun_out/ab_env_$i.err').read()[-1500:])
PY
done
done
