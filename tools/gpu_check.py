"""GPU diagnostic sweep (run under gpurun): every conv configuration on both kernels against the numpy oracle,
continuing past failures and writing a JSON report to gpurun_out/gpu_check.json.  Not a test: a microscope."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import planer_oracle as oracle          # noqa: E402  (checker only)
import planer_b200 as planer            # noqa: E402
from planer_b200 import ops             # noqa: E402
from planer_b200 import backend as B    # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def conv_case(n, cin, h, w, cout, k, stride=1, pad=None, dil=1, dtype='float16', algo=ops.ALGO_AUTO, bias=True,
              bn=False, res=False, act=0, res_after=False, seed=0):
    pad = (k // 2) * dil if pad is None else pad
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, cin, h, w)).astype(dtype)
    K = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(dtype)
    b = (rng.standard_normal(cout) * 0.1).astype(dtype) if bias else None
    ref = oracle.conv2d(x.astype(np.float32), K.astype(np.float32), None if b is None else b.astype(np.float32),
                        1, (stride, stride), (dil, dil), (pad,) * 4)
    bk = bb = None
    if bn:
        bk = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        bb = (rng.standard_normal(cout) * 0.1).astype(np.float32)
        ref = ref * bk.reshape(1, -1, 1, 1) + bb.reshape(1, -1, 1, 1)
    r = None
    if res:
        r = rng.standard_normal(ref.shape).astype(dtype)
    if r is not None and not res_after:
        ref = ref + r.astype(np.float32)
    if act == 1:
        ref = np.maximum(ref, 0)
    elif act == 2:
        ref = np.where(ref > 0, ref, ref * 0.1)
    if r is not None and res_after:
        ref = ref + r.astype(np.float32)

    dt = np.dtype(dtype)
    cpad = cin if (dt != np.float16 or cin % 16 == 0) else (cin + 15) // 16 * 16
    xd = B.to_nhwc(B.asarray(x), dt, cpad)
    Kd = B.asarray(K)
    wp = ops.pack_weight(Kd, cpad, dt)
    y = B.empty(ref.shape, dt, 'nhwc')
    scale = shift = None
    if b is not None or bn:
        scale, shift = ops.fold_affine(None if b is None else B.asarray(b.astype(np.float32)),
                                       None if bk is None else B.asarray(bk), None if bb is None else B.asarray(bb), cout)
    rd = B.to_nhwc(B.asarray(r)) if r is not None else None
    picked = ops.conv2d_algo(xd, y, k, k, (stride, stride), (dil, dil), (pad,) * 4)
    ops.conv2d_into(xd, wp, y, k, k, (stride, stride), (dil, dil), (pad,) * 4, 1, scale, shift, rd, act, 0.1, algo,
                    res_after)
    B.synchronize()
    out = y.get()
    return rel(out, ref), picked


CASES = [
    # (label, kwargs)   -- tcgen05 path first, smallest first
    ('1x1 64->64 8x8 n1', dict(n=1, cin=64, h=8, w=8, cout=64, k=1, bias=False)),
    ('1x1 64->64 16x16 n2 (M=512)', dict(n=2, cin=64, h=16, w=16, cout=64, k=1, bias=False)),
    ('1x1 64->128', dict(n=2, cin=64, h=16, w=16, cout=128, k=1)),
    ('1x1 128->256', dict(n=2, cin=128, h=16, w=16, cout=256, k=1)),
    ('1x1 256->512 (2 n-tiles)', dict(n=2, cin=256, h=16, w=16, cout=512, k=1)),
    ('3x3 64->64 p1 14x14', dict(n=2, cin=64, h=14, w=14, cout=64, k=3)),
    ('3x3 64->64 p1 56x56 n4', dict(n=4, cin=64, h=56, w=56, cout=64, k=3)),
    ('3x3 64->64 p0', dict(n=2, cin=64, h=14, w=14, cout=64, k=3, pad=0)),
    ('3x3 s2 64->128 28x28', dict(n=2, cin=64, h=28, w=28, cout=128, k=3, stride=2)),
    ('3x3 s2 odd 17x19', dict(n=2, cin=64, h=17, w=19, cout=64, k=3, stride=2)),
    ('3x3 d2 p2', dict(n=2, cin=64, h=14, w=14, cout=64, k=3, dil=2)),
    ('1x1 s2 64->128', dict(n=2, cin=64, h=28, w=28, cout=128, k=1, stride=2, pad=0)),
    ('3x3 128->128 28x28', dict(n=2, cin=128, h=28, w=28, cout=128, k=3)),
    ('3x3 256->256 14x14', dict(n=2, cin=256, h=14, w=14, cout=256, k=3)),
    ('3x3 512->512 7x7 n4', dict(n=4, cin=512, h=7, w=7, cout=512, k=3)),
    ('3x3 cin32 (64B swizzle)', dict(n=2, cin=32, h=14, w=14, cout=64, k=3)),
    ('3x3 cin16 (32B swizzle)', dict(n=2, cin=16, h=14, w=14, cout=32, k=3)),
    ('3x3 cin96 (64B swizzle, 3 chunks)', dict(n=2, cin=96, h=10, w=10, cout=64, k=3)),
    ('7x7 s2 p3 cin3->16 pad', dict(n=2, cin=3, h=64, w=64, cout=64, k=7, stride=2, pad=3, bias=False)),
    ('1x1 64->255 (odd Cout)', dict(n=1, cin=64, h=13, w=13, cout=255, k=1)),
    ('1x1 512->1000 (dense-like)', dict(n=1, cin=512, h=1, w=4, cout=1000, k=1)),
    ('epilogue bn+relu', dict(n=2, cin=64, h=14, w=14, cout=64, k=3, bn=True, act=1)),
    ('epilogue bn+res+relu', dict(n=2, cin=64, h=14, w=14, cout=64, k=3, bn=True, res=True, act=1)),
    ('epilogue bn+leaky+res_after', dict(n=2, cin=64, h=14, w=14, cout=64, k=3, bn=True, res=True, act=2, res_after=True)),
    ('big 3x3 64->64 56x56 n32', dict(n=32, cin=64, h=56, w=56, cout=64, k=3, bn=True, act=1)),
    ('big 3x3 256->256 14x14 n32', dict(n=32, cin=256, h=14, w=14, cout=256, k=3)),
]


def run_cases(indices, algos):
    planer.core(planer.b200)
    rows = []
    for i in indices:
        label, kw = CASES[i]
        for algo_name in algos:
            algo = ops.ALGO_DIRECT if algo_name == 'direct' else ops.ALGO_AUTO
            t0 = time.time()
            try:
                err, picked = conv_case(algo=algo, **kw)
                status = 'ok' if err < 1e-2 else 'MISMATCH'
            except Exception as e:
                err, picked, status = None, None, 'ERROR: %s' % (str(e).splitlines()[0][:300],)
            row = {'i': i, 'case': label, 'algo': algo_name, 'picked': picked, 'rel_err': err, 'status': status,
                   'sec': round(time.time() - t0, 3)}
            rows.append(row)
            print('ROW ' + json.dumps(row), flush=True)
    return rows


def main():
    import argparse
    import subprocess
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='each', choices=['each', 'direct', 'auto', 'f32'])
    ap.add_argument('--index', type=int, default=-1)
    a = ap.parse_args()
    idx = list(range(len(CASES))) if a.index < 0 else [a.index]
    if a.mode in ('direct', 'auto'):
        run_cases(idx, [a.mode])
        return 0
    if a.mode == 'f32':
        planer.core(planer.b200)
        for label, kw in CASES[:12:3]:
            try:
                err, picked = conv_case(algo=ops.ALGO_AUTO, **dict(kw, dtype='float32'))
                row = {'case': label + ' [f32]', 'algo': 'auto', 'picked': picked, 'rel_err': err,
                       'status': 'ok' if err < 1e-3 else 'MISMATCH'}
            except Exception as e:
                row = {'case': label + ' [f32]', 'status': 'ERROR: %s' % str(e)[:300]}
            print('ROW ' + json.dumps(row), flush=True)
        return 0
    # each: the direct kernel and the fp32 cases in one child each, the tensor-core kernel one child PER CASE so
    # that a faulting launch cannot poison the rest of the sweep
    report = []

    def child(args, timeout):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + args, capture_output=True, text=True,
                               timeout=timeout)
            out, tail = r.stdout, (r.stderr or '')[-600:]
        except subprocess.TimeoutExpired as e:
            out, tail = (e.stdout or b'').decode() if isinstance(e.stdout, bytes) else (e.stdout or ''), 'TIMEOUT'
        rows = [json.loads(l[4:]) for l in out.splitlines() if l.startswith('ROW ')]
        return rows, tail

    rows, tail = child(['--mode', 'direct'], 900)
    report += rows
    print('direct: %d rows, %d ok %s' % (len(rows), sum(r['status'] == 'ok' for r in rows), tail if len(rows) < len(CASES) else ''), flush=True)
    rows, tail = child(['--mode', 'f32'], 300)
    report += rows
    for i in range(len(CASES)):
        rows, tail = child(['--mode', 'auto', '--index', str(i)], 240)
        if not rows:
            rows = [{'i': i, 'case': CASES[i][0], 'algo': 'auto', 'status': 'CRASH', 'stderr': tail}]
        report += rows
        print(json.dumps(rows[0]), flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'gpu_check.json'), 'w') as f:
        json.dump(report, f, indent=1)
    bad = [r for r in report if r['status'] != 'ok']
    for r in bad:
        print('BAD', json.dumps(r))
    print('SUMMARY: %d rows, %d not ok' % (len(report), len(bad)))
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
