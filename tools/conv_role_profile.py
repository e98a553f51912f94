"""Where do the cycles of one tensor-core conv launch go, per warp role?  (debug; run under gpurun)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import planer_b200 as planer
from planer_b200 import ops, backend as B, _capi
planer.core(planer.b200)
lib, ctx = B.lib(), B.ctx()
def run(n, cin, h, w, cout, k, stride=1, res=False):
    rng = np.random.default_rng(0)
    x = B.to_nhwc(B.asarray(rng.standard_normal((n, cin, h, w)).astype(np.float16)))
    K = B.asarray((rng.standard_normal((cout, cin, k, k)) * 0.05).astype(np.float16))
    wp = ops.pack_weight(K, cin, np.float16)
    pad = k // 2
    oh = (h + 2 * pad - k) // stride + 1
    y = B.empty((n, cout, oh, oh), np.float16, 'nhwc')
    r = B.empty((n, cout, oh, oh), np.float16, 'nhwc') if res else None
    for _ in range(3):
        ops.conv2d_into(x, wp, y, k, k, (stride,) * 2, (1, 1), (pad,) * 4, residual=r, act=1)
    _capi.check(lib.plnr_debug_conv_profile(ctx, 1, None, 0))
    ops.conv2d_into(x, wp, y, k, k, (stride,) * 2, (1, 1), (pad,) * 4, residual=r, act=1)
    out = (C.c_int64 * (148 * 8))()
    _capi.check(lib.plnr_debug_conv_profile(ctx, 1, out, 148 * 8))
    a = np.array(out[:]).reshape(148, 8).astype(np.float64)
    m = a.mean(0)
    print('conv %dx%d %d->%d @%d n=%d s=%d res=%d' % (k, k, cin, cout, h, n, stride, res))
    print('  producer: wait_empty %8.0f of %8.0f cycles (%.0f%%)' % (m[0], m[1], 100 * m[0] / m[1]))
    print('  mma     : wait_full  %8.0f, wait_acc %8.0f of %8.0f (%.0f%% / %.0f%%)' % (m[2], m[3], m[4], 100 * m[2] / m[4], 100 * m[3] / m[4]))
    print('  epilogue: wait_acc   %8.0f of %8.0f (%.0f%%)' % (m[5], m[6], 100 * m[5] / m[6]), flush=True)
    _capi.check(lib.plnr_debug_conv_profile(ctx, 0, None, 0))
run(128, 64, 56, 56, 64, 3)
run(128, 64, 56, 56, 64, 3, res=True)
run(128, 128, 28, 28, 128, 3)
run(128, 256, 14, 14, 256, 3)
run(128, 256, 14, 14, 256, 3, res=True)
run(128, 512, 7, 7, 512, 3)
run(128, 64, 56, 56, 128, 1)
