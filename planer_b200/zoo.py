"""Synthetic model graphs in Planer's JSON IR.

The reference's ONNX importer cannot run here (no ``onnx`` package) and its zoo downloader needs
the network, so the graphs named by BASELINE.json's configs are authored directly in the IR that
``read_onnx`` would have produced (reference: planer/io.py:53-287 -- one layer per flow entry,
BatchNormalization pre-folded to ``<gamma>_invK`` / ``<gamma>_invB`` inits of shape (1,C,1,1)
with eps hard-coded to 1e-5 (io.py:76-91), Gemm -> 'dense' with a ``shp`` attr (io.py:110-111),
a trailing ``return`` layer (io.py:284-285), weights concatenated into one flat uint8 blob in
``inits`` order (io.py:286)).  Weights are seeded random numbers of the right shapes ("data":
"synthetic"); nothing here downloads or reads a checkpoint.

Every builder returns ``(model, blob)`` with ``model = {'input', 'inits', 'layers', 'flow'}``
and ``blob`` a 1-D uint8 numpy array; ``save_model`` writes the ``.json`` + ``.npy`` pair (or the
``.pla`` zip) that ``read_net`` loads (planer/io.py:8-34, 289-299).
"""
import io as _io
import json
import zipfile

import numpy as np


class _Builder:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.inits, self.layers, self.flow, self.blobs = [], [], [], []
        self.uid = 0

    def init(self, name, arr):
        arr = np.ascontiguousarray(arr)
        self.inits.append([name, list(arr.shape), str(arr.dtype)])
        self.blobs.append(arr.reshape(-1).view(np.uint8))
        return name

    def op(self, kind, attrs, ins, out=None, name=None):
        self.uid += 1
        name = name or '%s_%d' % (kind, self.uid)
        out = out or name + ':0'
        self.layers.append([name, kind, attrs])
        self.flow.append([ins[0] if len(ins) == 1 else list(ins), [name], out])
        return out

    def conv(self, x, cin, cout, k, stride=1, pad=None, dil=1, group=1, bias=False, name=None, std=None):
        pad = (k // 2) * dil if pad is None else pad
        fan = (cin // group) * k * k
        std = np.sqrt(2.0 / fan) if std is None else std
        name = name or 'conv_%d' % (self.uid + 1)
        w = (self.rng.standard_normal((cout, cin // group, k, k)) * std).astype(np.float32)
        ins = [x, self.init(name + '.weight', w)]
        if bias:
            ins.append(self.init(name + '.bias', (self.rng.standard_normal(cout) * 0.1).astype(np.float32)))
        attrs = {'group': group, 'strides': [stride, stride], 'dilations': [dil, dil], 'pads': [pad] * 4}
        return self.op('conv', attrs, ins, name=name)

    def bn(self, x, c, name=None, gamma_scale=1.0):
        """Random (gamma, beta, mean, var) folded exactly like planer/io.py:76-91."""
        name = name or 'bn_%d' % (self.uid + 1)
        gamma = (self.rng.uniform(0.5, 1.5, c) * gamma_scale).astype(np.float32)
        beta = (self.rng.standard_normal(c) * 0.1).astype(np.float32)
        mean = (self.rng.standard_normal(c) * 0.1).astype(np.float32)
        var = self.rng.uniform(0.5, 1.5, c).astype(np.float32)
        v_inv = 1 / np.sqrt(var + 1e-5)
        inv_b = (-gamma * mean * v_inv + beta).reshape(1, -1, 1, 1)
        inv_k = (gamma * v_inv).reshape(1, -1, 1, 1)
        ins = [x, self.init(name + '.weight_invK', inv_k), self.init(name + '.weight_invB', inv_b)]
        return self.op('batchnorm', {}, ins, name=name)

    def finish(self, inputs, outputs):
        self.layers.append(['return', 'return', {}])
        self.flow.append([list(outputs), ['return'], 'plrst'])
        model = {'input': list(inputs), 'inits': self.inits, 'layers': self.layers, 'flow': self.flow}
        blob = np.concatenate(self.blobs) if self.blobs else np.zeros(0, np.uint8)
        return model, blob


def single_conv(cin=3, cout=64, k=3, stride=1, pad=0, dil=1, group=1, bias=True, seed=0):
    """BASELINE config 1: the README's ``Conv2d(3, 64, 3, 1)`` as a one-layer graph."""
    b = _Builder(seed)
    y = b.conv('x', cin, cout, k, stride, pad, dil, group, bias, name='conv')
    return b.finish(['x'], [y])


def stem_net(cout=64, k=7, pad=3, bias=False, bn=True, seed=0):
    """The first four layers of a ResNet: conv k x k / s2 -> (batchnorm) -> relu -> maxpool 3/s2/p1.  The pattern the
    fused first-layer kernel (csrc/stem_pool.cu) absorbs."""
    b = _Builder(seed)
    x = b.conv('x', 3, cout, k, 2, pad, bias=bias, name='conv1')
    if bn:
        x = b.bn(x, cout, 'bn1')
    x = b.op('relu', {}, [x], name='relu')
    x = b.op('maxpool', {'w': [3, 3], 'pads': [1, 1, 1, 1], 'strides': [2, 2]}, [x], name='maxpool')
    return b.finish(['x'], [x])


def down_block(cin=64, cout=128, stride=2, seed=0):
    """conv3x3+relu, then one down-sampling BasicBlock: conv3x3/s -> bn -> relu -> conv3x3 -> bn, shortcut conv1x1/s -> bn,
    add, relu.  The pattern whose shortcut convolution the executor folds into the second conv's launch."""
    b = _Builder(seed)
    x = b.conv('x', cin, cin, 3, 1, 1, name='pre')
    x = b.op('relu', {}, [x], name='pre_relu')
    y = b.conv(x, cin, cout, 3, stride, 1, name='conv1')
    y = b.bn(y, cout, 'bn1')
    y = b.op('relu', {}, [y], name='relu1')
    y = b.conv(y, cout, cout, 3, 1, 1, name='conv2')
    y = b.bn(y, cout, 'bn2')
    idt = b.conv(x, cin, cout, 1, stride, 0, name='downsample.0')
    idt = b.bn(idt, cout, 'downsample.1')
    y = b.op('add', {}, [y, idt], name='add')
    y = b.op('relu', {}, [y], name='relu2')
    return b.finish(['x'], [y])


def decoder_net(seed=0, width=32):
    """A small encoder-decoder of the kind the reference's README demos run (U-Net / GAN style, readme.md:105-108), in the
    real IR: conv+bn+relu -> averagepool(2) -> conv+relu -> convtranspose(4, s2, p1)+bias -> relu -> concat with the skip ->
    convtranspose(3, s1, p1) -> sigmoid.  Covers the two SURVEY 8f rank-2 operators (planer/layer.py:28-34, :74-75)."""
    b = _Builder(seed)
    w = width
    x = b.conv('x', 3, w, 3, name='enc1')
    x = b.bn(x, w, name='enc1.bn')
    e1 = b.op('relu', {}, [x], name='enc1.relu')
    x = b.op('averagepool', {'w': [2, 2], 'pads': [0, 0, 0, 0], 'strides': [2, 2]}, [e1], name='pool')
    x = b.conv(x, w, 2 * w, 3, name='enc2', bias=True)
    x = b.op('relu', {}, [x], name='enc2.relu')
    kt = (b.rng.standard_normal((2 * w, w, 4, 4)) * np.sqrt(2.0 / (2 * w * 4))).astype(np.float32)
    ins = [x, b.init('up.weight', kt), b.init('up.bias', (b.rng.standard_normal(w) * 0.1).astype(np.float32))]
    x = b.op('convtranspose', {'strides': [2, 2], 'dilations': [1, 1], 'pads': [1, 1, 1, 1], 'output_padding': [0, 0],
                               'group': 1}, ins, name='up')
    x = b.op('relu', {}, [x], name='up.relu')
    x = b.op('concat', {'axis': 1}, [e1, x], name='cat')
    ko = (b.rng.standard_normal((2 * w, 3, 3, 3)) * np.sqrt(2.0 / (2 * w * 9))).astype(np.float32)
    x = b.op('convtranspose', {'strides': [1, 1], 'dilations': [1, 1], 'pads': [1, 1, 1, 1], 'output_padding': [0, 0],
                               'group': 1}, [x, b.init('out.weight', ko)], name='out')
    y = b.op('sigmoid', {}, [x], name='out.sigmoid')
    return b.finish(['x'], [y])


def upsample_net(seed=0, width=16):
    """conv+relu -> bilinear x2 (UpSample, mode 'linear') -> conv+relu -> Resize nearest x2 (ONNX default modes) -> conv ->
    softmax over channels: the interpolation operators of SURVEY 8f rank 2 in the real IR (planer/layer.py:80-88, :141-146)."""
    b = _Builder(seed)
    w = width
    x = b.conv('x', 3, w, 3, name='c1', bias=True)
    x = b.op('relu', {}, [x], name='c1.relu')
    x = b.op('upsample', {'mode': 'linear'}, [x, b.init('up1.scales', np.array([1, 1, 2, 2], np.float32))], name='up1')
    x = b.conv(x, w, w, 3, name='c2')
    x = b.op('relu', {}, [x], name='c2.relu')
    x = b.op('resize', {'mode': 'nearest', 'coordinate_transformation_mode': 'half_pixel', 'nearest_mode': 'round_prefer_floor'},
             [x, b.init('up2.roi', np.zeros(0, np.float32)), b.init('up2.scales', np.array([1, 1, 2, 2], np.float32))], name='up2')
    x = b.conv(x, w, 5, 1, name='c3', bias=True)
    y = b.op('softmax', {'axis': 1}, [x], name='prob')
    return b.finish(['x'], [y])


def readme_net(seed=0):
    """The README's CustomNet in the real IR: conv+relu chained in one flow,
    maxpool(2), upsample(x2, nearest), concat(axis=1)+sigmoid chained, return."""
    rng = np.random.default_rng(seed)
    K = (rng.standard_normal((64, 3, 3, 3)) * np.sqrt(2.0 / 27)).astype(np.float32)
    B = (rng.standard_normal(64) * 0.1).astype(np.float32)
    S = np.array([1, 1, 2, 2], np.float32)
    model = {
        'input': ['x'],
        'inits': [['K', [64, 3, 3, 3], 'float32'], ['B', [64], 'float32'], ['S', [4], 'float32']],
        'layers': [['conv', 'conv', {'group': 1, 'strides': [1, 1], 'dilations': [1, 1], 'pads': [1, 1, 1, 1]}],
                   ['relu', 'relu', {}],
                   ['pool', 'maxpool', {'w': [2, 2], 'pads': [0, 0, 0, 0], 'strides': [2, 2]}],
                   ['up', 'upsample', {'mode': 'nearest'}],
                   ['concat', 'concat', {'axis': 1}], ['sigmoid', 'sigmoid', {}], ['return', 'return', {}]],
        'flow': [[['x', 'K', 'B'], ['conv', 'relu'], 'a'], ['a', ['pool'], 'p'], [['p', 'S'], ['up'], 'y'],
                 [['a', 'y'], ['concat', 'sigmoid'], 'z'], [['z'], ['return'], 'plrst']],
    }
    blob = np.concatenate([a.reshape(-1).view(np.uint8) for a in (K, B, S)])
    return model, blob


def resnet18(seed=0, num_classes=1000):
    """ResNet-18 @224: conv1 7x7/s2/p3 -> bn -> relu -> maxpool 3/s2/p1 ->
    8 BasicBlocks (conv-bn-relu-conv-bn [+ 1x1/s2 downsample conv-bn] -> add -> relu) ->
    gap -> flatten -> dense -> return.  70 layers, 62 inits, BN kept as separate layers."""
    b = _Builder(seed)
    x = b.conv('x', 3, 64, 7, 2, 3, name='conv1')
    x = b.bn(x, 64, 'bn1')
    x = b.op('relu', {}, [x], name='relu')
    x = b.op('maxpool', {'w': [3, 3], 'pads': [1, 1, 1, 1], 'strides': [2, 2]}, [x], name='maxpool')
    cin = 64
    for li, (cout, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], 1):
        for bi in range(2):
            s = stride if bi == 0 else 1
            p = 'layer%d.%d.' % (li, bi)
            idt = x
            y = b.conv(x, cin, cout, 3, s, 1, name=p + 'conv1')
            y = b.bn(y, cout, p + 'bn1')
            y = b.op('relu', {}, [y], name=p + 'relu1')
            y = b.conv(y, cout, cout, 3, 1, 1, name=p + 'conv2')
            y = b.bn(y, cout, p + 'bn2')
            if s != 1 or cin != cout:
                idt = b.conv(x, cin, cout, 1, s, 0, name=p + 'downsample.0')
                idt = b.bn(idt, cout, p + 'downsample.1')
            y = b.op('add', {}, [y, idt], name=p + 'add')
            x = b.op('relu', {}, [y], name=p + 'relu2')
            cin = cout
    x = b.op('gap', {}, [x], name='avgpool')
    x = b.op('flatten', {}, [x], name='flatten')
    w = (b.rng.standard_normal((num_classes, 512)) * np.sqrt(1.0 / 512)).astype(np.float32)
    bias = (b.rng.standard_normal(num_classes) * 0.1).astype(np.float32)
    x = b.op('dense', {'shp': [512, num_classes]}, [x, b.init('fc.weight', w), b.init('fc.bias', bias)], name='fc')
    return b.finish(['x'], [x])


def yolov3(seed=0, num_out=255, width=1.0):
    """Synthetic YOLOv3 (Darknet-53 + 3 heads): every conv is
    ``conv(no bias, pad=k//2) -> batchnorm -> leakyrelu(0.1)`` except the three linear 1x1 head
    outputs (with bias).  ``width`` scales channel counts (tests use a narrow copy)."""
    b = _Builder(seed)
    ch = lambda c: max(8, int(c * width))

    def cbl(x, cin, cout, k, stride=1, gamma_scale=1.0):
        y = b.conv(x, cin, cout, k, stride)
        y = b.bn(y, cout, gamma_scale=gamma_scale)
        return b.op('leakyrelu', {'alpha': 0.1}, [y])

    x = cbl('x', 3, ch(32), 3)
    cin, routes = ch(32), []
    for n_res, c in [(1, 64), (2, 128), (8, 256), (8, 512), (4, 1024)]:
        c = ch(c)
        x = cbl(x, cin, c, 3, 2)
        for _ in range(n_res):
            y = cbl(x, c, c // 2, 1)
            y = cbl(y, c // 2, c, 3, gamma_scale=0.3)     # damped residual branch: 23 shortcut adds would
            x = b.op('add', {}, [x, y])                   # otherwise grow activations past the fp16 range
        cin = c
        routes.append((x, c))

    def head(x, cin, c):
        for _ in range(2):
            x = cbl(x, cin, c, 1)
            x = cbl(x, c, c * 2, 3)
            cin = c * 2
        branch = cbl(x, cin, c, 1)
        y = cbl(branch, c, c * 2, 3)
        out = b.conv(y, c * 2, num_out, 1, 1, 0, bias=True)
        return branch, out

    outs = []
    x, cin = routes[4]
    branch, o = head(x, cin, ch(512))
    outs.append(o)
    for ri, c in ((3, ch(256)), (2, ch(128))):
        y = cbl(branch, c * 2, c, 1)
        b.uid += 1
        up = 'upsample_%d' % b.uid
        y = b.op('upsample', {'mode': 'nearest'},
                 [y, b.init(up + '.scales', np.array([1, 1, 2, 2], np.float32))], name=up)
        rx, rc = routes[ri]
        x = b.op('concat', {'axis': 1}, [y, rx])
        branch, o = head(x, c + rc, c)
        outs.append(o)
    return b.finish(['x'], outs)


def conv_flops(model, input_shape):
    """Algorithmic conv+dense FLOPs of one forward: 2*N*Co*(C/g)*kh*kw*oh*ow, true K."""
    from .plan import infer_shapes
    shapes = infer_shapes(model, {model['input'][0]: tuple(input_shape)})
    return sum(s['flops'] for s in shapes['nodes'] if 'flops' in s)


def save_model(path, model, blob, pla=False):
    """Write ``path.json`` + ``path.npy`` (or ``path.pla``), the formats of planer/io.py:8-24,289-299."""
    import os
    if pla:
        base = os.path.split(path)[1]
        buf = _io.BytesIO()
        np.save(buf, blob)
        with zipfile.ZipFile(path + '.pla', 'w') as f:
            f.writestr(base + '.json', json.dumps(model))
            f.writestr(base + '.npy', buf.getvalue())
    else:
        with open(path + '.json', 'w') as f:
            json.dump(model, f)
        np.save(path + '.npy', blob)
