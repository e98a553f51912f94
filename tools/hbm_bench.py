"""Achieved HBM bandwidth of the streaming companions of the conv path (maxpool, upsample, concat, sigmoid, relu, add,
batchnorm scale/shift, layout transposes), fp16, at sizes well above the 126 MB L2; run under gpurun.

    python tools/hbm_bench.py [out.json]

Algorithmic bytes = (elements read + elements written) * 2; time = CUDA events around 10 back-to-back launches
on rotating buffers; peak = MEASURED_PEAKS.json hbm_gbs (copy bandwidth) when present.
"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import planer_b200 as planer
from planer_b200 import ops, backend as B, _capi
planer.core(planer.b200)
lib, ctx = B.lib(), B.ctx()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6650.0
if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']


def timed(fn, reps=10):
    for _ in range(3):
        fn(0)
    B.synchronize()
    a, b = C.c_void_p(), C.c_void_p()
    lib.plnr_event_create(C.byref(a)); lib.plnr_event_create(C.byref(b))
    lib.plnr_event_record(ctx, a)
    for i in range(reps):
        fn(i)
    lib.plnr_event_record(ctx, b)
    B.synchronize()
    ms = C.c_float(); lib.plnr_event_elapsed_ms(a, b, C.byref(ms))
    return ms.value / reps


def rnd(shape, layout='nhwc'):
    a = B.empty(shape, np.float16, layout)
    lib.plnr_memset(ctx, a.ptr, 0x3c, a.size * 2)
    return a


rows = []


def report(name, ref, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows.append(dict(kernel=name, replaces=ref, mbytes=nbytes / 1e6, us=ms * 1e3, gbs=gbs, frac_of_measured_copy_peak=gbs / peak))
    print('%-34s %8.1f MB %8.1f us %8.0f GB/s  %.2f of %.0f GB/s' % (name, nbytes / 1e6, ms * 1e3, gbs, gbs / peak, peak), flush=True)


NB = 3                                                  # rotating buffer sets: every launch streams fresh data
# maxpool 3x3/s2/p1 on the ResNet-18 stem output: 128 x 64 x 112 x 112 -> 56 x 56
xs = [rnd((128, 64, 112, 112)) for _ in range(NB)]; ys = [rnd((128, 64, 56, 56)) for _ in range(NB)]
report('maxpool_kernel 3x3/s2/p1', 'planer/util.py:79-95', (xs[0].size + ys[0].size) * 2,
       timed(lambda i: ops.maxpool_into(xs[i % NB], ys[i % NB], (3, 3), (1, 1, 1, 1), (2, 2))))
y2 = [rnd((128, 64, 56, 56)) for _ in range(NB)]
report('maxpool_kernel 2x2/s2 (README net)', 'planer/util.py:79-95', (xs[0].size + y2[0].size) * 2,
       timed(lambda i: ops.maxpool_into(xs[i % NB], y2[i % NB], (2, 2), (0, 0, 0, 0), (2, 2))))
# nearest upsample x2 (YOLOv3, batch 32 x 8 to exceed L2): 256 x 256 x 26 x 26 -> 52 x 52
xu = [rnd((256, 256, 26, 26)) for _ in range(NB)]; yu = [rnd((256, 256, 52, 52)) for _ in range(NB)]
report('upsample_kernel x2 nearest', 'planer/util.py:184-192', (xu[0].size + yu[0].size) * 2,
       timed(lambda i: ops.upsample_into(xu[i % NB], yu[i % NB], 2, 2)))
# concat along channels = two channel-slice copies into one buffer (256 + 512 @ 26x26, batch 256)
xa = [rnd((256, 256, 26, 26)) for _ in range(NB)]; xb = [rnd((256, 512, 26, 26)) for _ in range(NB)]
yc = [rnd((256, 768, 26, 26)) for _ in range(NB)]
def concat(i):
    ops.copy_channels(xa[i % NB], ops.channel_slice(yc[i % NB], 0, 256))
    ops.copy_channels(xb[i % NB], ops.channel_slice(yc[i % NB], 256, 512))
report('copy_channels_kernel (concat)', 'planer/layer.py:90-91', 2 * yc[0].size * 2, timed(concat))
# elementwise family on 128 x 64 x 112 x 112 (205 MB per tensor)
e_in = [rnd((128, 64, 112, 112)) for _ in range(NB)]; e_in2 = rnd((128, 64, 112, 112)); e_out = [rnd((128, 64, 112, 112)) for _ in range(NB)]
n_e = e_in[0].size
report('eltwise_kernel sigmoid', 'planer/layer.py:61-64', 2 * n_e * 2, timed(lambda i: ops.eltwise(ops.EW_SIGMOID, e_in[i % NB], e_out[i % NB])))
report('eltwise_kernel relu (in place)', 'planer/layer.py:44-46', 2 * n_e * 2, timed(lambda i: ops.eltwise(ops.EW_RELU, e_in[i % NB], e_in[i % NB])))
report('eltwise_kernel add', 'planer/layer.py:93-95', 3 * n_e * 2, timed(lambda i: ops.eltwise(ops.EW_ADD, e_in[i % NB], e_out[i % NB], p0=e_in2)))
k = B.asarray(np.random.default_rng(0).uniform(0.5, 1.5, 64).astype(np.float16)); b = B.asarray(np.zeros(64, np.float16))
report('eltwise_kernel scale_shift (BN)', 'planer/layer.py:125-127', 2 * n_e * 2,
       timed(lambda i: ops.eltwise(ops.EW_SCALE_SHIFT, e_in[i % NB], e_out[i % NB], p0=k, p1=b)))
# graph-exit layout transpose NHWC -> NCHW
flat = [B.empty((128, 64, 112, 112), np.float16) for _ in range(NB)]
report('nhwc_to_nchw_kernel', 'planer/net.py:100', 2 * n_e * 2, timed(lambda i: ops.nhwc_to_nchw_into(e_in[i % NB], flat[i % NB])))
# ---- the operators next to the path ----
ya = [rnd((128, 64, 56, 56)) for _ in range(NB)]
report('avgpool_kernel 2x2/s2', 'planer/util.py:97-100', (xs[0].size + ya[0].size) * 2,
       timed(lambda i: ops.avgpool_into(xs[i % NB], ya[i % NB], (2, 2), (0, 0, 0, 0), (2, 2))))
report('avgpool_kernel 3x3/s2/p1', 'planer/util.py:97-100', (xs[0].size + ya[0].size) * 2,
       timed(lambda i: ops.avgpool_into(xs[i % NB], ya[i % NB], (3, 3), (1, 1, 1, 1), (2, 2))))
report('upsample_linear_kernel x2', 'planer/util.py:133-153', (xu[0].size + yu[0].size) * 2,
       timed(lambda i: ops.upsample_linear_into(xu[i % NB], yu[i % NB], 2, 2)))
# zero stuffing of ConvTranspose2d(k4, s2, p1): 256 x 256 x 26 x 26 -> 55 x 55
yz = [rnd((256, 256, 55, 55)) for _ in range(NB)]
report('zero_stuff_kernel s2 (convtranspose)', 'planer/layer.py:32-33', (xu[0].size + yz[0].size) * 2,
       timed(lambda i: ops.zero_stuff_into(xu[i % NB], yz[i % NB], 2, 2, (2, 2))))
report('unary2_kernel clip', 'planer/layer.py:247-251', 2 * n_e * 2, timed(lambda i: ops.unary2(ops.EW_CLIP, e_in[i % NB], e_out[i % NB], 0, 6)))
report('unary2_kernel hardsigmoid', 'planer/layer.py:66-69', 2 * n_e * 2,
       timed(lambda i: ops.unary2(ops.EW_HARDSIGMOID, e_in[i % NB], e_out[i % NB], 0.2, 0.5)))
report('softmax_kernel (64 channels per pixel)', 'planer/layer.py:141-146', 2 * n_e * 2, timed(lambda i: ops.softmax_into(e_in[i % NB], e_out[i % NB])))
if len(sys.argv) > 1:
    json.dump(dict(peak_gbs=peak, rows=rows), open(sys.argv[1], 'w'), indent=1)
