"""YOLOv3-416 fp16 forward (BASELINE configs[3]: batch 32, synthetic Darknet-53 + 3 heads): images/s, TFLOP/s and the
per-kernel-kind time split (CUDA events, un-graphed pass); run under gpurun."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import planer_b200 as planer
from planer_b200 import zoo, backend as B
planer.core(planer.b200)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model, blob = zoo.yolov3(0)
net = planer.from_model(model, blob, half=True)
rng = np.random.default_rng(1)
xs = [B.asarray(rng.standard_normal((n, 3, 416, 416)).astype(np.float16)) for _ in range(3)]
ex = net.executor([(n, 3, 416, 416)])
for i in range(3):
    net.forward(xs[i % 3])
B.synchronize()
lib, ctx = B.lib(), B.ctx()
a, b = C.c_void_p(), C.c_void_p()
lib.plnr_event_create(C.byref(a)); lib.plnr_event_create(C.byref(b))
steps = 20
lib.plnr_event_record(ctx, a)
for i in range(steps):
    net.forward(xs[i % 3])
lib.plnr_event_record(ctx, b)
B.synchronize()
ms = C.c_float(); lib.plnr_event_elapsed_ms(a, b, C.byref(ms))
per = ms.value / steps
flops = ex.plan.flops
print(json.dumps({'model': 'yolov3-416 fp16', 'batch': n, 'ms_per_step': per, 'images_per_s': n / per * 1e3,
                  'tflops': flops / per / 1e9, 'gflop_per_step': flops / 1e9, 'launches_per_step': len(ex.launches)}))
# per-kind split
ex._load_inputs([xs[0]])
kinds, per_step = {}, []
for fn, kind, name in zip(ex.launches, ex.kinds, ex.names):
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.plnr_event_create(C.byref(e0)); lib.plnr_event_create(C.byref(e1))
    lib.plnr_event_record(ctx, e0); fn(); lib.plnr_event_record(ctx, e1)
    B.synchronize()
    t = C.c_float(); lib.plnr_event_elapsed_ms(e0, e1, C.byref(t))
    kinds[kind] = kinds.get(kind, 0.0) + t.value
    per_step.append((t.value, kind, name))
print('eager per-kind ms (includes launch gaps):', {k: round(v, 3) for k, v in kinds.items()})
fl = {nd.name: nd.flops for nd in ex.plan.nodes}
print('slowest launches (ms, TFLOP/s, fused step):')
for t, kind, name in sorted(per_step, reverse=True)[:14]:
    f = fl.get(name.split('+')[0], 0)
    print('  %.3f  %6.0f  %s %s' % (t, f / t / 1e9 if t > 0 else 0, kind, name[:70]))
