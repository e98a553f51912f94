"""Time the fused first-layer kernel alone (CUDA events, rotating inputs > L2); run under gpurun."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import planer_b200 as planer
from planer_b200 import ops, backend as B, _capi, zoo
planer.core(planer.b200)
lib, ctx = B.lib(), B.ctx()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model, blob = zoo.stem_net(64, 7, 3, False, True, seed=0)
net = planer.from_model(model, blob, half=True)
rng = np.random.default_rng(0)
xs = [B.asarray(rng.standard_normal((n, 3, 224, 224)).astype(np.float16)) for _ in range(4)]
ex = net.executor([(n, 3, 224, 224)])
run = list(ex.fused_stems.values())[0]['run']
for i in range(5):
    run(xs[i % 4])
B.synchronize()
reps = 20
a, b = C.c_void_p(), C.c_void_p()
lib.plnr_event_create(C.byref(a)); lib.plnr_event_create(C.byref(b))
lib.plnr_event_record(ctx, a)
for i in range(reps):
    run(xs[i % 4])
lib.plnr_event_record(ctx, b)
B.synchronize()
ms = C.c_float(); lib.plnr_event_elapsed_ms(a, b, C.byref(ms))
print('stem_pool n=%d: %.1f us per launch (%d back-to-back launches), PLNR_STEM_BAND=%s' % (n, ms.value / reps * 1e3, reps, os.environ.get('PLNR_STEM_BAND', '7')))

_capi.check(lib.plnr_debug_conv_profile(ctx, 1, None, 0))
run(xs[0])
out = (C.c_int64 * 2048)()
_capi.check(lib.plnr_debug_conv_profile(ctx, 1, out, 2048))
a = np.array(out[:148 * 8]).reshape(148, 8).astype(np.float64)
a = a[a[:, 4] > 0]
m = a.mean(0)
print('roles (mean cycles over CTAs): producer0 wait_empty %.0f of %.0f | mma wait_A %.0f wait_acc %.0f of %.0f | epilogue wait_acc %.0f barrier %.0f of %.0f'
      % (m[0], m[1], m[2], m[3], m[4], m[5], m[7], m[6]))
_capi.check(lib.plnr_debug_conv_profile(ctx, 0, None, 0))
