"""Same-box comparator: the op the reference runs on a GPU is cuDNN's convolution (planer/util.py:66-77 ``conv_dnn``,
selected at planer/layer.py:24 when cupy + cudnn are installed).  This tool times cuDNN 9 (through torch: fp16,
channels_last, ``cudnn.benchmark`` picks its best algorithm) on every conv shape of the bench config and on the whole
network, next to OUR per-kernel CUDA-event table from the same process and box, and writes a markdown table.

    python tools/cudnn_compare.py [--config resnet18|yolov3] [--batch N] [--out profiles/r02_vs_cudnn.md]

cuDNN is the COMPARATOR only: nothing in planer_b200/ calls it.  Times are CUDA events over `--reps` back-to-back launches
after warm-up (cuDNN) and the best of 3 un-graphed event pairs (ours, includes ~2 us launch gap per kernel); the whole-net
rows are graph replays on both sides.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_torch(fn, reps, warm=10):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def torch_net_from_model(model, blob, half=True):
    """The zoo IR as a plain torch callable (conv with BatchNorm folded into weight + bias, the way cuDNN users deploy):
    returns fn(x) built from torch.nn.functional ops only.  Supports what resnet18 / yolov3 use."""
    import torch
    import torch.nn.functional as F
    inits, s = {}, 0
    for name, shape, dt in model['inits']:
        dt = np.dtype(dt)
        n = int(np.prod(shape)) * dt.itemsize
        inits[name] = np.frombuffer(blob[s:s + n].tobytes(), dtype=dt).reshape(shape)
        s += n
    layers = {n: (k, a) for n, k, a in model['layers']}
    dev = torch.device('cuda')
    tdt = torch.float16 if half else torch.float32
    cache = {}

    def T(name):
        if name not in cache:
            cache[name] = torch.from_numpy(inits[name].astype(np.float32)).to(dev)
        return cache[name]

    # fold conv -> batchnorm pairs
    flow = [(xs if isinstance(xs, list) else [xs], ls if isinstance(ls, list) else [ls], y) for xs, ls, y in model['flow']]
    prog = []
    for xs, ls, y in flow:
        for j, l in enumerate(ls):
            kind, attrs = layers[l]
            prog.append([kind, attrs, list(xs) if j == 0 else [y], y])
    folded, skip = [], set()
    for i, (kind, attrs, ins, out) in enumerate(prog):
        if i in skip:
            continue
        if kind == 'conv':
            w = T(ins[1]).clone()
            b = T(ins[2]).clone() if len(ins) > 2 else torch.zeros(w.shape[0], device=dev)
            j = i + 1
            if j < len(prog) and prog[j][0] == 'batchnorm' and prog[j][2][0] == out:
                k_, b_ = T(prog[j][2][1]).reshape(-1), T(prog[j][2][2]).reshape(-1)
                w = w * k_.reshape(-1, 1, 1, 1)
                b = b * k_ + b_
                out = prog[j][3]
                skip.add(j)
            w = w.to(tdt).contiguous(memory_format=torch.channels_last)
            folded.append(('conv', dict(attrs), [ins[0]], out, (w, b.to(tdt))))
        else:
            folded.append((kind, dict(attrs), ins, out, None))

    def run(x):
        env = {model['input'][0]: x}
        res = None
        for kind, a, ins, out, par in folded:
            if kind == 'conv':
                p = a.get('pads', [0, 0, 0, 0])
                env[out] = F.conv2d(env[ins[0]], par[0], par[1], stride=tuple(a.get('strides', (1, 1))), padding=(p[0], p[1]),
                                    dilation=tuple(a.get('dilations', (1, 1))), groups=a.get('group', 1))
            elif kind == 'relu':
                env[out] = F.relu(env[ins[0]])
            elif kind == 'leakyrelu':
                env[out] = F.leaky_relu(env[ins[0]], a.get('alpha', 0.2))
            elif kind == 'maxpool':
                env[out] = F.max_pool2d(env[ins[0]], tuple(a['w']), tuple(a['strides']), (a['pads'][0], a['pads'][1]))
            elif kind == 'add':
                env[out] = env[ins[0]] + env[ins[1]]
            elif kind == 'gap':
                env[out] = env[ins[0]].mean((2, 3), keepdim=True)
            elif kind == 'flatten':
                env[out] = env[ins[0]].flatten(1)
            elif kind == 'dense':
                env[out] = F.linear(env[ins[0]], T(ins[1]).to(tdt), T(ins[2]).to(tdt))
            elif kind == 'upsample':
                env[out] = F.interpolate(env[ins[0]], scale_factor=2, mode='nearest')
            elif kind == 'concat':
                env[out] = torch.cat([env[i] for i in ins], 1)
            elif kind == 'sigmoid':
                env[out] = torch.sigmoid(env[ins[0]])
            elif kind == 'return':
                res = tuple(env[i] for i in ins)
            else:
                raise NotImplementedError(kind)
        return res
    return run


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='resnet18', choices=['resnet18', 'yolov3'])
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--reps', type=int, default=50)
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'r02_vs_cudnn.md'))
    args = ap.parse_args()
    import torch
    import torch.nn.functional as F
    import planer_b200 as planer
    from planer_b200 import zoo, backend as B
    import bench
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    planer.core(planer.b200)
    hw = 224 if args.config == 'resnet18' else 416
    batch = args.batch or (128 if args.config == 'resnet18' else 32)
    model, blob = zoo.resnet18(0) if args.config == 'resnet18' else zoo.yolov3(0)
    net = planer.from_model(model, blob, half=True)
    shape = (batch, 3, hw, hw)
    rng = np.random.default_rng(0)
    xh = rng.standard_normal(shape).astype(np.float16)
    xd = B.asarray(xh)
    ex = net.executor([shape])
    for _ in range(5):
        net.forward(xd)
    B.synchronize()
    table = bench.per_kernel_profile(net, xd)
    ours_step = bench.time_graph_steps(net, [xd], 30)

    # ---- per-layer: cuDNN conv (+bias) alone on the same shapes, vs our fused kernel (conv+bn+add+act) ----
    from planer_b200 import plan as P
    vals = ex.plan.values
    rows = []
    ours_by_name = {r['name'].split('+')[0]: r for r in table}
    dev = torch.device('cuda')
    for st in ex.plan.steps:
        if st.op != 'conv':
            continue
        xs, ys, ks = vals[st.ins[0]].shape, vals[st.out].shape, vals[st.w].shape
        a = st.attrs
        x = torch.randn(xs, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        w = torch.randn(ks, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        b = torch.randn(ks[0], device=dev, dtype=torch.float16)
        fn = lambda: F.conv2d(x, w, b, stride=a['strides'], padding=(a['pads'][0], a['pads'][1]), dilation=a['dilations'],
                              groups=a['group'])
        ms = time_torch(fn, args.reps)
        flops = 2.0 * ys[0] * ys[1] * ys[2] * ys[3] * ks[1] * ks[2] * ks[3]
        mine = ours_by_name.get(st.name)
        extra = 0.0
        if st.shortcut is not None:                      # our launch also computes the 1x1 shortcut conv: time cuDNN's too
            sh = st.shortcut[2]
            x2s, k2s = vals[sh.ins[0]].shape, vals[sh.w].shape
            x2 = torch.randn(x2s, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
            w2 = torch.randn(k2s, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
            extra = time_torch(lambda: F.conv2d(x2, w2, None, stride=sh.attrs['strides']), args.reps)
            flops += 2.0 * ys[0] * ys[1] * ys[2] * ys[3] * k2s[1]
        rows.append(dict(name='+'.join(st.fused)[:60], x=xs, k=ks, stride=a['strides'][0], gflop=flops / 1e9,
                         cudnn_ms=ms + extra, ours_ms=mine['ms'] if mine else None))
        del x, w, b
    # ---- whole net: torch (cuDNN convs, BN folded, eager ops captured in a CUDA graph) vs our graph ----
    tnet = torch_net_from_model(model, blob, half=True)
    xt = torch.from_numpy(xh).to(dev).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(3):
            yt = tnet(xt)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            tnet(xt)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            yt = tnet(xt)
        torch_step = time_torch(g.replay, 30, warm=5)
    y_ours = net(xh)
    y_ours = y_ours if isinstance(y_ours, tuple) else (y_ours,)
    agree = max(float(np.abs(a.astype(np.float32) - t.float().cpu().numpy()).max() / max(np.abs(t.float().cpu().numpy()).max(), 1e-30))
                for a, t in zip(y_ours, yt))

    lines = ['# cuDNN 9 (torch %s, cudnn %s) vs planer_b200 on the same B200, same process -- %s fp16 batch %d'
             % (torch.__version__, torch.backends.cudnn.version(), args.config, batch), '',
             'cuDNN is what the reference runs on a GPU (`planer/util.py:66-77`).  Per-layer rows: cuDNN = convolution + bias only '
             '(fp16, channels_last, `cudnn.benchmark` autotuned, CUDA events over %d back-to-back launches); ours = the fused launch '
             '(conv + BatchNorm + residual add + activation, best of 3 un-graphed CUDA-event pairs, which include the launch gap).  '
             'A row with a fused 1x1 shortcut counts cuDNN\'s shortcut conv too.' % args.reps, '',
             '| layer (our fused launch) | input | filter | s | GFLOP | cuDNN ms | cuDNN TFLOP/s | ours ms | ours TFLOP/s | ours / cuDNN speed |',
             '|---|---|---|---|---|---|---|---|---|---|']
    for r in rows:
        o = r['ours_ms']
        lines.append('| %s | %s | %s | %d | %.1f | %.4f | %.0f | %s | %s | %s |' % (
            r['name'], 'x'.join(map(str, r['x'])), 'x'.join(map(str, r['k'])), r['stride'], r['gflop'], r['cudnn_ms'],
            r['gflop'] / r['cudnn_ms'], '%.4f' % o if o else '-', '%.0f' % (r['gflop'] / o) if o else '-',
            '%.2fx' % (r['cudnn_ms'] / o) if o else '-'))
    tot_c = sum(r['cudnn_ms'] for r in rows)
    tot_o = sum(r['ours_ms'] for r in rows if r['ours_ms'])
    lines += ['', 'Sum over the conv layers: cuDNN %.3f ms (convs only), ours %.3f ms (fused launches, un-graphed).' % (tot_c, tot_o), '',
              '## Whole network (graph replay on both sides, inputs resident)', '',
              '| | ms / step | images/s |', '|---|---|---|',
              '| torch + cuDNN (BN folded into the convs, relu / add / pool as torch ops, one CUDA graph) | %.3f | %.0f |' % (torch_step, batch / torch_step * 1e3),
              '| planer_b200 (first-layer kernel + one CUDA graph) | %.3f | %.0f |' % (ours_step, batch / ours_step * 1e3),
              '', 'Outputs agree to %.2e range-relative (same folded weights, fp16 both sides).' % agree, '']
    with open(args.out, 'w') as f:
        f.write('\n'.join(lines))
    print('\n'.join(lines))
    print(json.dumps({'config': args.config, 'batch': batch, 'cudnn_step_ms': torch_step, 'ours_step_ms': ours_step,
                      'conv_sum_cudnn_ms': tot_c, 'conv_sum_ours_ms': tot_o, 'agree': agree}))


if __name__ == '__main__':
    main()
