#!/bin/bash
# A/B of two builds of the library on ONE box, YOLOv3 config: tools/ab_yolo.sh <pytest -k selection> <libA.so> <libB.so>
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "$1" 2>&1 | tail -3
for tag in A B A B; do
  lib=$2; [ $tag = B ] && lib=$3
  PLNR_LIB=$lib python bench.py --config yolov3 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/aby_$tag.line 2>/dev/null
  python - <<PY
import json
d = json.load(open('gpurun_out/aby_$tag.line'))
fam = {f['family']: round(f['ms'] * 1e3, 1) for f in d['roofline']['families']}
print('$tag', '$lib', round(d['value']), 'img/s', round(d['ms_per_step'] * 1e3, 1), 'us/step |', fam, flush=True)
PY
done
