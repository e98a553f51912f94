#!/bin/bash
# Round-2 GPU call 1: parity tests, stride-2 shift A/B, full bench lines (ResNet-18 + YOLOv3), cuDNN comparator.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -25 gpurun_out/r2_pytest.log
for v in 1 0 1 0; do
  PLNR_NO_SHIFT_S2=$v timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-e2e --dump gpurun_out/r2_pk_nos2_$v.json \
      > gpurun_out/r2_bench_nos2_$v.json 2> gpurun_out/r2_bench_nos2_$v.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_bench_nos2_$v.json')); t = json.load(open('gpurun_out/r2_pk_nos2_$v.json'))['table']
    print('NO_SHIFT_S2=$v', round(d['value']), 'img/s', round(d['ms_per_step'] * 1e3, 1), 'us/step |', ' '.join('%.0f' % (r['ms'] * 1e3) for r in t), flush=True)
except Exception as e:
    print('NO_SHIFT_S2=$v failed', e); print(open('gpurun_out/r2_bench_nos2_$v.err').read()[-1500:])
PY
done
timeout 900 python bench.py --steps 200 --dump gpurun_out/r2_pk.json > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
timeout 600 python bench.py --config yolov3 --steps 30 --no-cpu-baseline --dump gpurun_out/r2_pk_yolo.json > gpurun_out/r2_bench_yolo.json 2> gpurun_out/r2_bench_yolo.err
cat gpurun_out/r2_bench_yolo.json; tail -5 gpurun_out/r2_bench_yolo.err
timeout 600 python tools/cudnn_compare.py --out gpurun_out/r2_vs_cudnn.md > gpurun_out/r2_cudnn.log 2>&1
tail -40 gpurun_out/r2_cudnn.log
