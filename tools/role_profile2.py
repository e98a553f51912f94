"""Per-role cycle accounting of one tensor-core conv launch (debug; run under gpurun).

For each configuration: the launch time (CUDA events, best of 5) and, from the kernel's own counters, how long the
TMA producer waited for free slots, the MMA issuer for operands / a free accumulator, the epilogue for a finished
accumulator -- averaged over CTAs, as a share of the role's total.
    python tools/role_profile2.py [keys]
"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import planer_b200 as planer
from planer_b200 import ops, backend as B, _capi
planer.core(planer.b200)
lib, ctx = B.lib(), B.ctx()


def run(name, n, cin, h, w, cout, k, stride=1, res=False):
    rng = np.random.default_rng(0)
    x = B.to_nhwc(B.asarray(rng.standard_normal((n, cin, h, w)).astype(np.float16)))
    K = B.asarray((rng.standard_normal((cout, cin, k, k)) * 0.05).astype(np.float16))
    wp = ops.pack_weight(K, cin, np.float16)
    pad = k // 2
    pads = (pad,) * 4
    oh = (h + 2 * pad - k) // stride + 1
    y = B.empty((n, cout, oh, oh), np.float16, 'nhwc')
    r = B.empty((n, cout, oh, oh), np.float16, 'nhwc') if res else None
    call = lambda: ops.conv2d_into(x, wp, y, k, k, (stride,) * 2, (1, 1), pads, residual=r, act=1)
    for _ in range(3):
        call()
    best = 1e9
    for _ in range(5):
        a, b = C.c_void_p(), C.c_void_p()
        lib.plnr_event_create(C.byref(a)); lib.plnr_event_create(C.byref(b))
        lib.plnr_event_record(ctx, a); call(); lib.plnr_event_record(ctx, b)
        B.synchronize()
        ms = C.c_float(); lib.plnr_event_elapsed_ms(a, b, C.byref(ms)); best = min(best, ms.value)
    _capi.check(lib.plnr_debug_conv_profile(ctx, 1, None, 0))
    call()
    out = (C.c_int64 * 2048)()
    _capi.check(lib.plnr_debug_conv_profile(ctx, 1, out, 2048))
    a = np.array(out[:148 * 8]).reshape(148, 8).astype(np.float64)
    a = a[a[:, 4] > 0]
    if not len(a):                                   # a kernel without role counters (conv_pw.cu)
        print('%-28s %.4f ms  %6.0f TFLOP/s (kernel: %s, no role counters)' % (name, best, 2.0 * n * oh * oh * cout * cin * k * k / 1e9 / best, B.last_kernel()), flush=True)
        _capi.check(lib.plnr_debug_conv_profile(ctx, 0, None, 0))
        return
    m = a.mean(0)
    gflop = 2.0 * n * oh * oh * cout * cin * k * k / 1e9
    print('%-28s %.4f ms  %6.0f TFLOP/s | producer wait %3.0f%% of %7.0f | mma wait_full %3.0f%% wait_acc %3.0f%% of %7.0f | '
          'epilogue wait %3.0f%% of %7.0f' % (name, best, gflop / best, 100 * m[0] / max(m[1], 1), m[1], 100 * m[2] / max(m[4], 1),
                                             100 * m[3] / max(m[4], 1), m[4], 100 * m[5] / max(m[6], 1), m[6]), flush=True)
    worst = int(np.argmax(a[:, 6]))
    print('      epilogue role total per CTA: min %.0f  mean %.0f  max %.0f (CTA row %d: mma total %.0f, tfull wait %.0f) | mma total min %.0f max %.0f'
          % (a[:, 6].min(), a[:, 6].mean(), a[:, 6].max(), worst, a[worst, 4], a[worst, 5], a[:, 4].min(), a[:, 4].max()), flush=True)
    if os.environ.get('ROLE_DUMP'):
        print('      per-CTA epilogue totals:', ' '.join('%.0f' % (v / 1000) for v in a[:, 6]))
        print('      per-CTA mma totals:     ', ' '.join('%.0f' % (v / 1000) for v in a[:, 4]))
    st = [int(v) for v in out[1400:1408]]
    print('      CTA0 stamps (clk since kernel entry): setup done %d | first A landed %d | mma loop done %d | epilogue done %d | '
          'final barrier %d | dealloc %d | setup->dealloc %.1f us (globaltimer)' % (tuple(st[1:7]) + ((st[7] - st[0]) / 1e3,)), flush=True)
    es = [int(v) for v in out[1408:1424]]
    print('      last epilogue of CTA0 warp 4 (clk since entry): accumulator ready %d | then (tmem_ld done, chunk done) x chunks: %s'
          % (es[0], ' '.join('%d' % (v - es[0]) for v in es[1:] if v)), flush=True)
    _capi.check(lib.plnr_debug_conv_profile(ctx, 0, None, 0))


cfgs = {
    'a': ('layer1 64->64 @56', 128, 64, 56, 56, 64, 3, 1, False),
    'b': ('layer1 64->64 @56 +res', 128, 64, 56, 56, 64, 3, 1, True),
    'c': ('layer2 128->128 @28', 128, 128, 28, 28, 128, 3, 1, False),
    'd': ('layer2 128->128 @28 +res', 128, 128, 28, 28, 128, 3, 1, True),
    'e': ('layer3 256->256 @14', 128, 256, 14, 14, 256, 3, 1, False),
    'f': ('layer4 512->512 @7', 128, 512, 7, 7, 512, 3, 1, False),
    'g': ('layer2.0 64->128 s2', 128, 64, 56, 56, 128, 3, 2, False),
    'h': ('layer3.0 128->256 s2', 128, 128, 28, 28, 256, 3, 2, False),
    'i': ('down 64->128 1x1 s2', 128, 64, 56, 56, 128, 1, 2, False),
    'l': ('layer4.0 256->512 s2', 128, 256, 14, 14, 512, 3, 2, False),
    'j': ('128->128 @14 x512 (layer2 FLOPs)', 512, 128, 14, 14, 128, 3, 1, False),
    'k': ('64->64 @28 x512 (layer1 FLOPs)', 512, 64, 28, 28, 64, 3, 1, False),
    # YOLOv3-416 batch 32, the early (HBM-heavy) layers
    'y': ('yolo 32->64 3x3 @208 +res', 32, 32, 208, 208, 64, 3, 1, True),
    'z': ('yolo 32->64 3x3 s2 @416', 32, 32, 416, 416, 64, 3, 2, False),
    'w': ('yolo 64->32 1x1 @208', 32, 64, 208, 208, 32, 1, 1, False),
    'v': ('yolo 64->128 3x3 s2 @208', 32, 64, 208, 208, 128, 3, 2, False),
    'u': ('yolo 128->64 1x1 @104', 32, 128, 104, 104, 64, 1, 1, False),
}
for key in (sys.argv[1] if len(sys.argv) > 1 else 'abcdefghi'):
    run(*cfgs[key])
