"""Where do the cycles of one tensor-core conv launch go, per warp role?  (debug; run under gpurun)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import planer_b200 as planer
from planer_b200 import ops, backend as B, _capi
planer.core(planer.b200)
lib, ctx = B.lib(), B.ctx()
def run(n, cin, h, w, cout, k, stride=1, res=False, kw=None, pads=None):
    rng = np.random.default_rng(0)
    zero = os.environ.get('PLNR_ZERO_INPUT') == '1'
    x = B.to_nhwc(B.asarray((rng.standard_normal((n, cin, h, w)) * (0 if zero else 1)).astype(np.float16)))
    kw = k if kw is None else kw
    K = B.asarray((rng.standard_normal((cout, cin, k, kw)) * 0.05).astype(np.float16))
    wp = ops.pack_weight(K, cin, np.float16)
    pad = k // 2
    pads = (pad,) * 4 if pads is None else pads
    oh = (h + pads[0] + pads[2] - k) // stride + 1
    ow = (w + pads[1] + pads[3] - kw) // stride + 1
    y = B.empty((n, cout, oh, ow), np.float16, 'nhwc')
    r = B.empty((n, cout, oh, ow), np.float16, 'nhwc') if res else None
    for _ in range(3):
        ops.conv2d_into(x, wp, y, k, kw, (stride,) * 2, (1, 1), pads, residual=r, act=1)
    _capi.check(lib.plnr_debug_conv_profile(ctx, 1, None, 0))
    ops.conv2d_into(x, wp, y, k, kw, (stride,) * 2, (1, 1), pads, residual=r, act=1)
    out = (C.c_int64 * 2048)()
    _capi.check(lib.plnr_debug_conv_profile(ctx, 1, out, 2048))
    a = np.array(out[:148 * 8]).reshape(148, 8).astype(np.float64)
    tr = np.array(out[1200:1200 + 160]).reshape(40, 4)
    m = a.mean(0)
    print('conv %dx%d %d->%d @%d n=%d s=%d res=%d  [CG=%s EPI=%s ZERO=%s]' % (k, k, cin, cout, h, n, stride, res, os.environ.get('PLNR_CTA_GROUP','2'), os.environ.get('PLNR_DEBUG_EPI','0'), os.environ.get('PLNR_ZERO_INPUT','0')))
    print('  producer: wait_empty %8.0f of %8.0f cycles (%.0f%%)' % (m[0], m[1], 100 * m[0] / m[1]))
    print('  mma     : wait_full  %8.0f, wait_acc %8.0f of %8.0f (%.0f%% / %.0f%%)' % (m[2], m[3], m[4], 100 * m[2] / m[4], 100 * m[3] / m[4]))
    print('  epilogue: wait_acc   %8.0f of %8.0f (%.0f%%)' % (m[5], m[6], 100 * m[5] / m[6]), flush=True)
    cg = 1 if os.environ.get('PLNR_CTA_GROUP') == '1' else 2
    mt = -(-(n * oh * ow) // (128 * cg)); ntile = min(256, -(-cout // 32) * 32); nn = -(-cout // ntile)
    tiles = mt * nn; unitsn = 148 // cg; per = -(-tiles // unitsn)
    stages_per_cta = per * k * kw * max(1, cin // 64)
    print('  max tiles/CTA %d -> %d stages; producer max %8.0f cycles -> %.0f cycles/stage' % (per, stages_per_cta, a[:, 1].max(), a[:, 1].max() / stages_per_cta))
    mt_ = a[:, 4]; mt_ = mt_[mt_ > 0]
    print('  MMA-warp total per CTA: min %d  p25 %d  median %d  p75 %d  max %d   (n=%d)' % (mt_.min(), np.percentile(mt_, 25), np.median(mt_), np.percentile(mt_, 75), mt_.max(), mt_.size))
    print('  slowest CTAs:', np.argsort(-a[:, 4])[:12].tolist(), ' fastest:', np.argsort(a[:, 4] + (a[:, 4] == 0) * 1e12)[:8].tolist())
    print('  CTA0 MMA-warp trace per tile [start, after tempty, first full, end] (cycles):')
    for i in range(min(per, 8)): print('    tile %d: %s  dur=%d' % (i, tr[i].tolist(), tr[i][3] - tr[i][0]))
    _capi.check(lib.plnr_debug_conv_profile(ctx, 0, None, 0))
import sys
cfgs = {'s': (128, 64, 112, 112, 64, 4, 1, False, 1, (2, 0, 1, 0)), 'a': (128, 64, 56, 56, 256, 3), 'b': (128, 256, 14, 14, 256, 3), 'c': (32, 256, 56, 56, 256, 3), 'd': (128, 64, 56, 56, 64, 3),
        'e': (128, 64, 14, 14, 256, 3), 'f': (8, 256, 56, 56, 256, 3)}
for k in (sys.argv[1] if len(sys.argv) > 1 else 'abcdef'):
    run(*cfgs[k])
