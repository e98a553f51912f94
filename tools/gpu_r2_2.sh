#!/bin/bash
# Round-2 GPU call: full GPU test suite, fp32 bench lines, cold start with / without the pack cache, YOLO bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log
for b in 1 32; do
  for small in 1 0; do
    PLNR_DIRECT_SMALL=$small timeout 600 python bench.py --dtype f32 --batch $b --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2_f32_b${b}_small$small.json 2> gpurun_out/r2_f32_b${b}_small$small.err
    python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_f32_b${b}_small$small.json'))
    print('fp32 batch $b PLNR_DIRECT_SMALL=$small: %.1f img/s %.3f ms/step, %.1f TFLOP/s = %.3f of FFMA peak, parity %.2e' % (d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['checks']['parity']['rel_err']))
except Exception as e:
    print('fp32 batch $b failed', e); print(open('gpurun_out/r2_f32_b${b}_small$small.err').read()[-1500:])
PY
  done
done
timeout 600 python tools/cold_start.py 128 2>&1 | tail -8 | tee gpurun_out/r2_cold_start.md
timeout 600 python bench.py --config yolov3 --steps 30 --no-cpu-baseline > gpurun_out/r2_bench_yolo.json 2> gpurun_out/r2_bench_yolo.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_bench_yolo.json'))
print('yolo value %.0f  e2e %.0f  views %s  fp16_host %.0f' % (d['value'], d['e2e']['value'], d['e2e']['result_views'], d['e2e']['fp16_host']['value']))
PY
