"""Seeded test cases shared by oracle/gen_golden.py (reference run), the oracle tests (CPU) and the
GPU parity tests.  A case is regenerated from its name alone, so fixtures store only outputs."""
import zlib

import numpy as np

from planer_b200 import zoo


def _rng(name):
    return np.random.default_rng(zlib.crc32(name.encode()))


def _x(rng, shape, dtype):
    return rng.standard_normal(shape).astype(dtype)


def _conv(name, xs, ks, bias=True, dtype='float32', **kw):
    rng = _rng(name)
    fan = ks[1] * ks[2] * ks[3]
    args = [_x(rng, xs, dtype), (rng.standard_normal(ks) * np.sqrt(2.0 / fan)).astype(dtype)]
    if bias:
        args.append((rng.standard_normal(ks[0]) * 0.1).astype(dtype))
    return 'conv', args, kw


def _convt(name, xs, ks, bias=True, dtype='float32', **kw):
    """ConvTranspose2d case: K is (C_in, C_out, kh, kw) (planer/layer.py:28-34)."""
    rng = _rng(name)
    fan = ks[0] * ks[2] * ks[3]
    args = [_x(rng, xs, dtype), (rng.standard_normal(ks) * np.sqrt(2.0 / fan)).astype(dtype)]
    if bias:
        args.append((rng.standard_normal(ks[1]) * 0.1).astype(dtype))
    return 'convtranspose', args, kw


# name -> builder;  every builder returns (op kind, positional args, attrs)
OP_CASES = {
    # BASELINE config 1: README Conv2d(3, 64, 3, 1) on 1x3x32x32, both paddings
    'c1_conv_p0': lambda n: _conv(n, (1, 3, 32, 32), (64, 3, 3, 3), strides=(1, 1), pads=(0, 0, 0, 0)),
    'c1_conv_p1': lambda n: _conv(n, (1, 3, 32, 32), (64, 3, 3, 3), strides=(1, 1), pads=(1, 1, 1, 1)),
    'conv_s2_p1': lambda n: _conv(n, (2, 16, 17, 19), (32, 16, 3, 3), strides=(2, 2), pads=(1, 1, 1, 1)),
    'conv_d2_p2': lambda n: _conv(n, (2, 16, 14, 14), (16, 16, 3, 3), dilations=(2, 2), pads=(2, 2, 2, 2)),
    'conv_7x7_s2_p3': lambda n: _conv(n, (2, 3, 64, 64), (64, 3, 7, 7), bias=False, strides=(2, 2), pads=(3, 3, 3, 3)),
    'conv_1x1_s2': lambda n: _conv(n, (2, 64, 14, 14), (128, 64, 1, 1), bias=False, strides=(2, 2)),
    'conv_g2': lambda n: _conv(n, (2, 16, 9, 9), (8, 8, 3, 3), group=2, pads=(1, 1, 1, 1)),
    'conv_dw8': lambda n: _conv(n, (1, 8, 10, 10), (8, 1, 3, 3), group=8, pads=(1, 1, 1, 1)),
    'conv_1x3_asym': lambda n: _conv(n, (1, 4, 8, 8), (6, 4, 1, 3), pads=(0, 1, 0, 1)),
    'conv_pad_tl_only': lambda n: _conv(n, (1, 8, 8, 8), (8, 8, 3, 3), pads=(1, 1, 0, 0)),
    'conv_64_64_3x3': lambda n: _conv(n, (2, 64, 14, 14), (64, 64, 3, 3), bias=False, pads=(1, 1, 1, 1)),
    'conv_64_64_3x3_f16': lambda n: _conv(n, (2, 64, 14, 14), (64, 64, 3, 3), dtype='float16', pads=(1, 1, 1, 1)),
    'conv_s2_p1_f16': lambda n: _conv(n, (2, 32, 17, 19), (64, 32, 3, 3), dtype='float16', strides=(2, 2), pads=(1, 1, 1, 1)),
    'conv_255_f16': lambda n: _conv(n, (1, 64, 13, 13), (255, 64, 1, 1), dtype='float16'),
    'dense': lambda n: ('dense', [_x(_rng(n), (4, 512), 'float32'),
                                 (_rng(n + 'w').standard_normal((1000, 512)) / 22.6).astype('float32'),
                                 _x(_rng(n + 'b'), (1000,), 'float32')], {}),
    'dense_f16': lambda n: ('dense', [_x(_rng(n), (4, 512), 'float16'),
                                     (_rng(n + 'w').standard_normal((1000, 512)) / 22.6).astype('float16'),
                                     _x(_rng(n + 'b'), (1000,), 'float16')], {}),
    'relu': lambda n: ('relu', [_x(_rng(n), (2, 8, 5, 7), 'float32')], {}),
    'relu_f16': lambda n: ('relu', [_x(_rng(n), (2, 8, 5, 7), 'float16')], {}),
    'leakyrelu': lambda n: ('leakyrelu', [_x(_rng(n), (2, 8, 5, 7), 'float32')], {'alpha': 0.1}),
    'leakyrelu_f16': lambda n: ('leakyrelu', [_x(_rng(n), (2, 8, 5, 7), 'float16')], {'alpha': 0.1}),
    'sigmoid': lambda n: ('sigmoid', [_x(_rng(n), (2, 8, 5, 7), 'float32') * 3], {}),
    'sigmoid_f16': lambda n: ('sigmoid', [_x(_rng(n), (2, 8, 5, 7), 'float16') * 3], {}),
    'add': lambda n: ('add', [_x(_rng(n), (2, 8, 5, 7), 'float32'), _x(_rng(n + '2'), (2, 8, 5, 7), 'float32')], {}),
    'batchnorm': lambda n: ('batchnorm', [_x(_rng(n), (2, 8, 5, 7), 'float32'),
                                         _rng(n + 'k').uniform(0.5, 1.5, (1, 8, 1, 1)).astype('float32'),
                                         _x(_rng(n + 'b'), (1, 8, 1, 1), 'float32')], {}),
    'batchnorm_f16': lambda n: ('batchnorm', [_x(_rng(n), (2, 8, 5, 7), 'float16'),
                                             _rng(n + 'k').uniform(0.5, 1.5, (1, 8, 1, 1)).astype('float16'),
                                             _x(_rng(n + 'b'), (1, 8, 1, 1), 'float16')], {}),
    # quirk Q2: zero padding + -1e4 floor on signed data
    'maxpool_k3s2p1_signed': lambda n: ('maxpool', [_x(_rng(n), (2, 8, 13, 15), 'float32') - 2.0],
                                        {'w': (3, 3), 'pads': (1, 1, 1, 1), 'strides': (2, 2)}),
    'maxpool_k2s2': lambda n: ('maxpool', [_x(_rng(n), (2, 8, 9, 11), 'float32')], {}),
    'maxpool_k3s2p1_f16': lambda n: ('maxpool', [_x(_rng(n), (2, 16, 12, 12), 'float16')],
                                     {'w': (3, 3), 'pads': (1, 1, 1, 1), 'strides': (2, 2)}),
    'maxpool_floor': lambda n: ('maxpool', [np.full((1, 8, 4, 4), -3e4, 'float32')], {}),
    'upsample_x2': lambda n: ('upsample', [_x(_rng(n), (2, 8, 5, 7), 'float32'), np.array([1, 1, 2, 2], 'float32')],
                              {'mode': 'nearest'}),
    'upsample_2x3_f16': lambda n: ('upsample', [_x(_rng(n), (1, 8, 4, 5), 'float16'), np.array([1, 1, 2, 3], 'float32')],
                                   {'mode': 'nearest'}),
    'upsample_linear_x2': lambda n: ('upsample', [_x(_rng(n), (2, 8, 5, 7), 'float32'), np.array([1, 1, 2, 2], 'float32')],
                                     {'mode': 'linear'}),
    'upsample_linear_3x2_f16': lambda n: ('upsample', [_x(_rng(n), (1, 16, 6, 5), 'float16'), np.array([1, 1, 3, 2], 'float32')],
                                          {'mode': 'linear'}),
    'resize_linear_x2': lambda n: ('resize', [_x(_rng(n), (1, 8, 9, 6), 'float32'), np.zeros(0, 'float32'),
                                              np.array([1, 1, 2, 2], 'float32')], {'mode': 'linear'}),
    'resize_linear_1p5': lambda n: ('resize', [_x(_rng(n), (2, 8, 9, 6), 'float32'), np.zeros(0, 'float32'),
                                               np.array([1, 1, 1.5, 2.5], 'float32')], {'mode': 'linear'}),
    # (fp32 only: with a float16 image the reference computes the coordinates in float16, where w - 1.001 rounds up to w - 1
    #  for w >= 5 and its own gather then indexes one past the edge, planer/util.py:204-208)
    'resize_linear_down': lambda n: ('resize', [_x(_rng(n), (1, 16, 12, 10), 'float32'), np.zeros(0, 'float32'),
                                                np.array([1, 1, 0.75, 1.3], 'float32')], {'mode': 'linear'}),
    'resize_nearest_x2': lambda n: ('resize', [_x(_rng(n), (2, 8, 4, 6), 'float32'), np.zeros(0, 'float32'),
                                               np.array([1, 1, 2, 2], 'float32')], {'mode': 'nearest'}),
    'resize_nearest_asym_floor_x3': lambda n: ('resize', [_x(_rng(n), (1, 8, 4, 5), 'float16'), np.zeros(0, 'float32'),
                                                          np.array([1, 1, 3, 3], 'float32')],
                                               {'mode': 'nearest', 'coordinate_transformation_mode': 'asymmetric',
                                                'nearest_mode': 'floor'}),
    'concat_c': lambda n: ('concat', [_x(_rng(n), (2, 8, 5, 7), 'float32'), _x(_rng(n + '2'), (2, 16, 5, 7), 'float32')],
                           {'axis': 1}),
    'gap': lambda n: ('gap', [_x(_rng(n), (2, 16, 7, 7), 'float32')], {}),
    'gap_f16': lambda n: ('gap', [_x(_rng(n), (2, 16, 7, 7), 'float16')], {}),
    'flatten': lambda n: ('flatten', [_x(_rng(n), (2, 16, 1, 1), 'float32')], {}),
    # AveragePool (zero padding, divisor always kh*kw) and ConvTranspose2d (planer/layer.py:28-34, :74-75)
    'averagepool_k2s2': lambda n: ('averagepool', [_x(_rng(n), (2, 8, 9, 11), 'float32')], {}),
    'averagepool_k3s2p1': lambda n: ('averagepool', [_x(_rng(n), (2, 8, 13, 15), 'float32') + 1.0],
                                     {'w': (3, 3), 'pads': (1, 1, 1, 1), 'strides': (2, 2)}),
    'averagepool_k3s1p1_f16': lambda n: ('averagepool', [_x(_rng(n), (2, 16, 12, 12), 'float16')],
                                         {'w': (3, 3), 'pads': (1, 1, 1, 1), 'strides': (1, 1)}),
    'hardsigmoid': lambda n: ('hardsigmoid', [_x(_rng(n), (2, 8, 5, 7), 'float32') * 3], {'alpha': 0.25, 'beta': 0.5}),
    'hardsigmoid_f16': lambda n: ('hardsigmoid', [_x(_rng(n), (2, 8, 5, 7), 'float16') * 3], {}),
    'clip': lambda n: ('clip', [_x(_rng(n), (2, 8, 5, 7), 'float32') * 3], {'min': -1.0, 'max': 2.0}),
    'clip_f16': lambda n: ('clip', [_x(_rng(n), (2, 8, 5, 7), 'float16') * 3], {'min': 0, 'max': 6}),
    'softmax_logits': lambda n: ('softmax', [_x(_rng(n), (5, 1000), 'float32') * 4], {'axis': -1}),
    'softmax_logits_f16': lambda n: ('softmax', [_x(_rng(n), (3, 77), 'float16') * 4], {'axis': 1}),
    'softmax_channels_64_f16': lambda n: ('softmax', [_x(_rng(n), (3, 64, 5, 7), 'float16') * 2], {'axis': 1}),
    'softmax_logits_16': lambda n: ('softmax', [_x(_rng(n), (37, 16), 'float32') * 3], {'axis': -1}),
    'softmax_channels': lambda n: ('softmax', [_x(_rng(n), (2, 21, 6, 5), 'float32') * 2], {'axis': 1}),
    'convtranspose_k4s2p1': lambda n: _convt(n, (2, 16, 7, 9), (16, 8, 4, 4), strides=(2, 2), pads=(1, 1, 1, 1)),
    'convtranspose_k3s2_outpad': lambda n: _convt(n, (1, 8, 6, 6), (8, 12, 3, 3), strides=(2, 2), pads=(1, 1, 1, 1),
                                                   output_padding=(1, 1)),
    'convtranspose_k2s2_nobias': lambda n: _convt(n, (2, 8, 5, 5), (8, 4, 2, 2), bias=False, strides=(2, 2)),
    'convtranspose_k3s1p1_d2': lambda n: _convt(n, (1, 4, 9, 9), (4, 6, 3, 3), strides=(1, 1), dilations=(2, 2), pads=(2, 2, 2, 2)),
    'convtranspose_64_f16': lambda n: _convt(n, (2, 64, 7, 7), (64, 64, 4, 4), dtype='float16', strides=(2, 2), pads=(1, 1, 1, 1)),
}


def make_case(name):
    return OP_CASES[name](name)


# ---------------------------------------------------------------------------------------------
# whole-graph cases: name -> (builder, input shape, dtype/half)
# ---------------------------------------------------------------------------------------------
BUILDERS = {
    'readme': lambda: zoo.readme_net(0),
    'resnet18': lambda: zoo.resnet18(0),
    'yolov3_quarter': lambda: zoo.yolov3(0, width=0.25),
    'yolov3': lambda: zoo.yolov3(0),
    'decoder': lambda: zoo.decoder_net(0),
    'upsample_net': lambda: zoo.upsample_net(0),
}
GRAPH_CASES = {
    'readme_f32': ('readme', (2, 3, 32, 32), False),
    'readme_f16': ('readme', (2, 3, 32, 32), True),
    'resnet18_f32_n1': ('resnet18', (1, 3, 224, 224), False),     # BASELINE config 2
    'resnet18_f32_n2': ('resnet18', (2, 3, 224, 224), False),
    'resnet18_f16_n1': ('resnet18', (1, 3, 224, 224), True),      # true numpy-fp16 path (slow: no BLAS)
    'resnet18_small_f32': ('resnet18', (3, 3, 64, 64), False),
    'yolov3_quarter_f32': ('yolov3_quarter', (2, 3, 96, 96), False),
    'yolov3_416_f32_n1': ('yolov3', (1, 3, 416, 416), False),     # BASELINE config 4 graph
    'decoder_f32': ('decoder', (2, 3, 32, 40), False),            # averagepool + convtranspose in a graph
    'decoder_f16': ('decoder', (2, 3, 32, 40), True),
    'upsample_net_f32': ('upsample_net', (2, 3, 12, 20), False),  # bilinear upsample, resize, channel softmax in a graph
    'upsample_net_f16': ('upsample_net', (2, 3, 12, 20), True),
}

_model_cache = {}


def get_model(key):
    if key not in _model_cache:
        _model_cache.clear()          # blobs are big (YOLOv3: 248 MB); keep one at a time
        _model_cache[key] = BUILDERS[key]()
    return _model_cache[key]


def make_graph_case(name):
    key, shape, half = GRAPH_CASES[name]
    model, blob = get_model(key)
    x = np.random.default_rng(1).standard_normal(shape).astype('float16' if half else 'float32')
    return model, blob, x, half


def sample(t, limit=20000):
    """Deterministic strided sample of a flattened tensor (whole tensor if small)."""
    flat = np.ascontiguousarray(t).reshape(-1)
    step = max(1, flat.size // limit)
    return flat[::step].copy()


def rel_err(y, ref):
    """Range-relative error  max|y - ref| / max|ref|  (SURVEY 8d primary metric)."""
    y, ref = np.asarray(y, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(y - ref).max() / max(np.abs(ref).max(), 1e-30))


# ---------------------------------------------------------------------------------------------
# sliding-window inference (planer/util.py:291-348): name -> (image shape, decorator kwargs, per-window function id)
# ---------------------------------------------------------------------------------------------
def tile_fn(kind):
    """Deterministic per-window functions standing in for a network: resolution-preserving, x2 up, /2 down, HWC -> HW."""
    if kind == 'same':
        return lambda im: im * np.float32(0.5) + np.float32(1)
    if kind == 'up2':
        return lambda im: np.repeat(np.repeat(im, 2, axis=0), 2, axis=1) * np.float32(0.25)
    if kind == 'down2':
        return lambda im: im[::2, ::2] + np.float32(3)
    if kind == 'gray':
        return lambda im: im.mean(axis=2)
    raise KeyError(kind)


TILE_CASES = {
    'grid_3x4': ((150, 210), dict(window=64, margin=0.25), 'same'),
    'rgb_margin_int': ((97, 131, 3), dict(window=48, margin=7), 'same'),
    'up2_rgb': ((80, 100, 3), dict(window=40, margin=0.2), 'up2'),
    'down2': ((128, 96), dict(window=64, margin=0.25), 'down2'),
    'sample_half': ((160, 120), dict(sample=0.5, window=32, margin=0.25), 'same'),
    'sample_tuple_rgb': ((90, 70, 3), dict(sample=(120, 100), window=64, margin=0.1), 'gray'),
    'single_window_glob': ((50, 70), dict(window=128, glob=32), 'same'),
    'single_window_sampled': ((50, 70, 3), dict(sample=1.5, window=256, glob=16), 'up2'),
}


def make_tile_case(name):
    shape, kw, kind = TILE_CASES[name]
    return _rng(name).standard_normal(shape).astype('float32') * 10, dict(kw), tile_fn(kind)


# ---------------------------------------------------------------------------------------------
# random DAGs of the hot-path operators (GPU parity sweeps, oracle-vs-reference sweep)
# ---------------------------------------------------------------------------------------------
def random_graph(seed):
    """A random DAG of the hot-path operators built with the zoo's IR builder: plain / strided convolutions with optional
    batchnorm and activation, ResNet- and Darknet-style residual blocks, pooling, nearest upsampling + concatenation with an
    earlier tensor, a second graph output taken from the middle."""
    from planer_b200.zoo import _Builder
    rng = np.random.default_rng(seed)
    b = _Builder(seed)
    cin = int(rng.choice([3, 16, 32]))
    if rng.integers(0, 3) == 0:                       # odd, non-square extents
        h, w = int(rng.integers(17, 50)), int(rng.integers(17, 50))
    else:
        h = w = int(rng.choice([32, 48, 64]))
    size = (h, w)
    x, c = 'x', cin
    seen = {}                      # (h, w) -> (name, channels) of an earlier tensor to concatenate with
    mid = None

    def act(y):
        kind = str(rng.choice(['relu', 'leakyrelu', 'sigmoid', 'none']))
        if kind == 'none':
            return y
        return b.op(kind, {'alpha': 0.1} if kind == 'leakyrelu' else {}, [y])

    for step in range(int(rng.integers(5, 10))):
        kind = str(rng.choice(['conv', 'conv', 'res', 'dark', 'pool', 'upcat', 'gconv', 'dconv', 'clip', 'avg', 'alias', 'hsig', 'convt', 'fork']))
        if kind == 'conv':
            co, k = int(rng.choice([16, 32, 64, 128] + ([256, 512] if h <= 16 else []))), int(rng.choice([1, 3, 3, 5]))
            s = 2 if (h >= 16 and rng.integers(0, 3) == 0) else 1
            y = b.conv(x, c, co, k, stride=s, bias=bool(rng.integers(0, 2)))
            if rng.integers(0, 2):
                y = b.bn(y, co)
            x, c = act(y), co
            if s == 2:
                h, w = (h + 1) // 2, (w + 1) // 2
        elif kind == 'gconv' and c % 4 == 0:          # grouped 3x3 convolution (CUDA-core kernel) + relu
            x = b.op('relu', {}, [b.conv(x, c, c, 3, group=4, bias=True)])
        elif kind == 'dconv':                         # dilated 3x3 convolution + batchnorm
            co = int(rng.choice([32, 64]))
            x, c = b.bn(b.conv(x, c, co, 3, dil=2), co), co
        elif kind == 'clip':
            x = b.op('clip', {}, [x, b.init('clip%d.min' % step, np.array(-0.5, np.float32)), b.init('clip%d.max' % step, np.array(2.0, np.float32))])
        elif kind == 'avg' and h >= 8:
            x = b.op('averagepool', {'w': [2, 2], 'pads': [0, 0, 0, 0], 'strides': [2, 2]}, [x]); h, w = h // 2, w // 2
        elif kind == 'alias':                         # the reference's ReLU works IN PLACE: x is mutated too, so this is 2 relu(x)
            y = b.op('relu', {}, [x])
            x = b.op('add', {}, [x, y])
        elif kind == 'hsig':
            x = b.op('hardsigmoid', {'alpha': 0.2, 'beta': 0.5}, [x])
        elif kind == 'convt' and h <= 16:             # ConvTranspose2d k4 / s2 / p1: doubles the extent
            co = int(rng.choice([16, 32]))
            wt = (rng.standard_normal((c, co, 4, 4)) * np.sqrt(1.0 / (c * 4))).astype(np.float32)
            x = b.op('convtranspose', {'strides': [2, 2], 'dilations': [1, 1], 'pads': [1, 1, 1, 1], 'output_padding': [0, 0], 'group': 1},
                     [x, b.init('ct%d.weight' % step, wt), b.init('ct%d.bias' % step, (rng.standard_normal(co) * 0.1).astype(np.float32))])
            c, h, w = co, 2 * h, 2 * w
        elif kind == 'fork':                          # two convolutions read the same tensor, their results are added
            y1 = b.op('relu', {}, [b.conv(x, c, c, 3, bias=True)])
            y2 = b.bn(b.conv(x, c, c, 1), c, gamma_scale=0.5)
            x = b.op('add', {}, [y1, y2])
        elif kind == 'res':        # relu(x + bn(conv3x3(x)))
            y = b.bn(b.conv(x, c, c, 3), c, gamma_scale=0.5)
            x = b.op('relu', {}, [b.op('add', {}, [y, x])])
        elif kind == 'dark':       # x + leaky(bn(conv3x3(leaky(bn(conv1x1(x))))))
            cm = max(8, c // 2)
            y = b.op('leakyrelu', {'alpha': 0.1}, [b.bn(b.conv(x, c, cm, 1), cm)])
            y = b.op('leakyrelu', {'alpha': 0.1}, [b.bn(b.conv(y, cm, c, 3), c, gamma_scale=0.5)])
            x = b.op('add', {}, [x, y])
        elif kind == 'pool' and h >= 8:
            seen[(h, w)] = (x, c)
            if rng.integers(0, 2):
                x = b.op('maxpool', {'w': [2, 2], 'pads': [0, 0, 0, 0], 'strides': [2, 2]}, [x]); h, w = h // 2, w // 2
            else:
                x = b.op('maxpool', {'w': [3, 3], 'pads': [1, 1, 1, 1], 'strides': [2, 2]}, [x]); h, w = (h + 1) // 2, (w + 1) // 2
        elif kind == 'upcat' and (2 * h, 2 * w) in seen:
            other, oc = seen[(2 * h, 2 * w)]
            y = b.op('upsample', {'mode': 'nearest'}, [x, b.init('up%d.scales' % step, np.array([1, 1, 2, 2], np.float32))])
            x, c = b.op('concat', {'axis': 1}, [y, other]), c + oc
            h, w = 2 * h, 2 * w
        if mid is None and step >= 2 and rng.integers(0, 2):
            mid = x
    if rng.integers(0, 3) == 0:                       # classifier tail: gap -> flatten -> dense
        nc = int(rng.choice([10, 100]))
        y = b.op('flatten', {}, [b.op('gap', {}, [x])])
        wd = (rng.standard_normal((nc, c)) * np.sqrt(1.0 / c)).astype(np.float32)
        x = b.op('dense', {'shp': [c, nc]}, [y, b.init('fc.weight', wd), b.init('fc.bias', (rng.standard_normal(nc) * 0.1).astype(np.float32))])
        mid = None if mid is None else mid
    outs = [x] if mid is None or mid == x else [x, mid]
    model, blob = b.finish(['x'], outs)
    return model, blob, cin, size
