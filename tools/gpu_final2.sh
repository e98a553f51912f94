#!/bin/bash
# Final bench lines of the shipped binary: ResNet-18 fp16 (default), YOLOv3, fp32 batch 1 / 32.
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 > gpurun_out/fin_resnet18_f16_b128.json 2> gpurun_out/fin_resnet18.err; echo "rc=$?"
python bench.py --config yolov3 --steps 50 --warmup 5 > gpurun_out/fin_yolov3_f16_b32.json 2> gpurun_out/fin_yolov3.err; echo "rc=$?"
python bench.py --dtype f32 --batch 32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/fin_resnet18_f32_b32.json 2> gpurun_out/fin_f32.err; echo "rc=$?"
python bench.py --dtype f32 --batch 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/fin_resnet18_f32_b1.json 2> gpurun_out/fin_f32b1.err; echo "rc=$?"
for f in gpurun_out/fin_*.json; do python - <<PY
import json
d = json.load(open('$f'))
print('$f', round(d['value'], 1), d['unit'], round(d['ms_per_step'], 4), 'ms | e2e', round(d['e2e']['value'], 1) if d.get('e2e') else None, '| frac', round(d['roofline']['frac'], 4), '| traffic', d['roofline'].get('traffic'), '| clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
