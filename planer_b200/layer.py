"""Operator table of the B200 path (Level B of SURVEY 8b).

Mirror of planer/layer.py for the hot-path operators: same names, same positional tensors + keyword
attributes, same aliasing behaviour (ReLU mutates and returns its input, planer/layer.py:44-46), same
error style (Python exceptions).  Each function launches hand-written sm_100a kernels through the C ABI on
``DeviceArray`` operands; operators outside the hot path are NOT provided -- ``layer_map[...]`` raises
``NotImplementedError`` naming the op instead of silently computing on the CPU.

``Net`` (net.py) normally does not call these one by one: it compiles the flow into a fused plan
(plan.py).  This eager table is what ``forward(debug=True)`` uses, what per-operator parity tests call,
and what ``planer_b200.install(planer)`` plugs into the *reference's* own ``Net``.
"""
import numpy as np

from . import _capi, ops
from . import backend as B
from .backend import DeviceArray


def wrap(f, layername='layer'):
    """planer/layer.py:6-13: turn an operator function into a Layer class."""
    class Layer:
        name = layername
        def __init__(self, **key): self.key = key
        def para(self): return self.key
        def forward(self, *x): return f(*x, **self.key)
        def __call__(self, *x): return self.forward(*x)
    return Layer


def _round_up(v, m):
    return (v + m - 1) // m * m


def _as_nhwc(x, dtype=None, pad16=False):
    """4-D operand -> internal pixel-major layout (graph inputs arrive flat NCHW)."""
    if not isinstance(x, DeviceArray):
        x = B.asarray(x)
    dtype = np.dtype(x.dtype if dtype is None else dtype)
    if x.layout == 'nhwc':
        return x if x.dtype == dtype else _cast_nhwc(x, dtype)
    c = x.shape[1]
    return B.to_nhwc(x, dtype, _round_up(c, 16) if pad16 and c % 16 else c)


def _cast_nhwc(x, dtype):
    flat = B.to_flat(x).astype(dtype)
    return B.to_nhwc(flat)


def _compute_dtype(*arrs):
    # a uint8 image takes the dtype of the float operand (numpy: uint8 x float16 -> float16)
    dts = {np.dtype(a.dtype) for a in arrs if a is not None and np.dtype(a.dtype) != np.uint8}
    return dts.pop() if len(dts) == 1 else np.dtype(np.float32)     # numpy would promote to fp32


_pack_cache = {}


def _packed(K, cin_pad, dtype):
    key = (K.ptr, K.shape, str(K.dtype), cin_pad, str(dtype))
    if key not in _pack_cache:
        if len(_pack_cache) > 4096: _pack_cache.clear()
        _pack_cache[key] = (ops.pack_weight(K, cin_pad, dtype), K)   # keep K alive: the key holds its address
    return _pack_cache[key][0]


def _packed_split(K):
    key = (K.ptr, K.shape, 'split')
    if key not in _pack_cache:
        if len(_pack_cache) > 4096: _pack_cache.clear()
        w16, meta = ops.pack_weight_split(K)
        _pack_cache[key] = (ops.split_weight(w16, meta, K.shape[1]), K)
    return _pack_cache[key][0]


def Conv2d(x, K, B_=None, group=1, strides=(1, 1), dilations=(1, 1), pads=(0, 0, 0, 0)):
    """planer/layer.py:22-26.  x (N,C,H,W); K (Co,C/g,kh,kw); B_ (Co) or None; pads = (top,left,bottom,right)."""
    dt = _compute_dtype(x, K)
    co, cg, kh, kw = K.shape
    pad16 = dt == np.float16 and group == 1
    xin = _as_nhwc(x, dt, pad16)
    cin = xin.shape[1]
    if cin != cg * group and not (group == 1 and cin >= cg):
        raise ValueError('Conv2d: input has %d channels, weight expects %d' % (x.shape[1], cg * group))
    y = B.empty(ops.conv_out_shape(x.shape, K.shape, strides, dilations, pads), dt, 'nhwc')
    if dt == np.float32 and K.dtype == np.float32 and group == 1 and cin == cg and ops.split_conv_enabled():
        wp = _packed_split(K)                # float32 on the tensor pipe (fp16 split operands, csrc/split_f32.cu)
    else:
        wp = _packed(K, cin // group, dt)
    scale = shift = None
    if B_ is not None:
        scale, shift = ops.fold_affine(B_, None, None, co)
    return ops.conv2d_into(xin, wp, y, kh, kw, strides, dilations, pads, group, scale, shift)


def Dense(x, K, B_, shp=None):
    """planer/layer.py:15-18: x @ K.T + B with K stored (out, in)."""
    dt = _compute_dtype(x, K)
    x2 = B.to_flat(x).astype(dt)
    if x2.ndim != 2:
        x2 = x2.reshape(x2.shape[0], -1)
    Kc = K.astype(dt)
    y = B.empty((x2.shape[0], K.shape[0]), dt)
    scale, shift = ops.fold_affine(B_, None, None, K.shape[0]) if B_ is not None else (None, None)
    return ops.dense_into(x2, Kc, y, scale, shift)


def _dense_rows(x):
    """Elementwise ops run on dense pixel rows; anything else is first made dense."""
    if not isinstance(x, DeviceArray):
        x = B.asarray(x)
    if x.layout == 'nhwc' and (x.ld != x.shape[1] or x.coff):
        y = B.empty(x.shape, x.dtype, 'nhwc')
        return ops.copy_channels(x, y)
    if x.layout == 'flat' and x.ndim == 4:
        return B.to_nhwc(x)
    return x


def _like(x):
    return B.empty(x.shape, x.dtype, x.layout)


def ReLU(x):
    """planer/layer.py:44-46: in place, returns the same object."""
    if x.layout == 'flat' and x.ndim == 4:          # NCHW graph input: elementwise ops are layout-agnostic
        flat = x.reshape(-1, 1)
        ops.eltwise(ops.EW_RELU, flat, flat)
        return x
    assert x.layout == 'flat' or (x.ld == x.shape[1] and x.coff == 0)
    ops.eltwise(ops.EW_RELU, x, x)
    return x


def LeakyReLU(x, alpha=0.2):
    """planer/layer.py:48-51 (new array)."""
    x = _dense_rows(x)
    return ops.eltwise(ops.EW_LEAKY, x, _like(x), alpha=alpha)


def Sigmoid(x):
    """planer/layer.py:61-64 (new array)."""
    x = _dense_rows(x)
    return ops.eltwise(ops.EW_SIGMOID, x, _like(x))


def HardSigmoid(x, alpha=0.2, beta=0.5):
    """planer/layer.py:66-69: max(min(x*alpha + beta, 1), 0)."""
    x = _dense_rows(x)
    return ops.unary2(ops.EW_HARDSIGMOID, x, _like(x), alpha, beta)


def Clip(x, min=0, max=1):
    """planer/layer.py:247-251: np.minimum(x, max, out=x); np.maximum(x, min, out=x) -- IN PLACE, returns its input (the
    numexpr branch of the reference, which allocates, is not the path the oracle pins)."""
    d = _dense_rows(x)           # x itself when it is dense; a dense copy of a strided view / NCHW graph input otherwise
    # bounds that arrive as inputs (constant inits of the graph, uploaded like every init) are read back: two scalars
    scalar = lambda v: float(np.asarray(v.get() if isinstance(v, DeviceArray) else v, np.float64).reshape(-1)[0])
    return ops.unary2(ops.EW_CLIP, d, d, scalar(min), scalar(max))


def Softmax(x, axis=-1):
    """planer/layer.py:141-146 along the channel axis of a 4-D tensor (axis 1) or the last axis of a 2-D one."""
    if not isinstance(x, DeviceArray):
        x = B.asarray(x)
    if x.ndim == 4 and axis in (1, -3):
        x = _dense_rows(x)
    elif x.ndim == 2 and axis in (1, -1):
        x = B.to_flat(x)
    else:
        raise NotImplementedError('Softmax: only the channel axis of 4-D tensors and the last axis of 2-D tensors are '
                                  'implemented (got %d-D, axis %d)' % (x.ndim, axis))
    return ops.softmax_into(x, _like(x))


def Add(x1, x2):
    """planer/layer.py:93-95 for equal-shape operands (the residual adds of the hot path)."""
    x1, x2 = _dense_rows(x1), _dense_rows(x2)
    if x1.shape != x2.shape or x1.dtype != x2.dtype:
        raise NotImplementedError('Add: broadcasting / mixed dtypes are outside the B200 hot path '
                                  '(%s %s vs %s %s)' % (x1.shape, x1.dtype, x2.shape, x2.dtype))
    return ops.eltwise(ops.EW_ADD, x1, _like(x1), p0=x2)


def BatchNorm(x, K, B_):
    """planer/layer.py:125-127: x*K + B with K, B of shape (1,C,1,1) (folded at import, planer/io.py:76-91)."""
    x = _dense_rows(x)
    dt = x.dtype
    return ops.eltwise(ops.EW_SCALE_SHIFT, x, _like(x), p0=K.astype(dt), p1=B_.astype(dt))


def Flatten(x):
    """planer/layer.py:59.  The flattened order is the reference's (c, h, w)."""
    if x.layout == 'nhwc' and x.shape[2] == x.shape[3] == 1 and x.ld == x.shape[1] and x.coff == 0:
        return DeviceArray(x.buf, (x.shape[0], x.shape[1]), x.dtype, 'flat', offset=x.offset)
    return B.to_flat(x).reshape(x.shape[0], -1)


def Maxpool(x, w=(2, 2), pads=(0, 0, 0, 0), strides=(2, 2)):
    """planer/layer.py:71-72 -> planer/util.py:79-95 (zero padding, -1e4 floor)."""
    if pads[2] > pads[0] or pads[3] > pads[1]:
        raise ValueError('Maxpool: bottom/right padding larger than top/left is undefined in the reference '
                         '(planer/util.py:4-10)')
    x = _as_nhwc(x)
    y = B.empty(ops.pool_out_shape(x.shape, w, pads, strides), x.dtype, 'nhwc')
    return ops.maxpool_into(x, y, w, pads, strides)


def AveragePool(x, w=(2, 2), pads=(0, 0, 0, 0), strides=(2, 2)):
    """planer/layer.py:74-75 -> planer/util.py:97-100 (zero padding, divisor kh*kw)."""
    if pads[2] > pads[0] or pads[3] > pads[1]:
        raise ValueError('AveragePool: bottom/right padding larger than top/left is undefined in the reference '
                         '(planer/util.py:4-10)')
    x = _as_nhwc(x)
    y = B.empty(ops.pool_out_shape(x.shape, w, pads, strides), x.dtype, 'nhwc')
    return ops.avgpool_into(x, y, w, pads, strides)


_flip_cache = {}


def _flipped(K):
    key = (K.ptr, K.shape, str(K.dtype))
    if key not in _flip_cache:
        if len(_flip_cache) > 1024: _flip_cache.clear()
        _flip_cache[key] = (ops.flip_weight(K), K)
    return _flip_cache[key][0]


def ConvTranspose2d(x, K, B_=None, strides=(2, 2), dilations=(1, 1), pads=(0, 0, 0, 0), output_padding=(0, 0), group=1):
    """planer/layer.py:28-34: zero-stuff the input, then an ordinary stride-1 convolution with the filter transposed and
    flipped -- the same tensor-core conv kernels.  K is (C_in, C_out, kh, kw)."""
    if group != 1:
        raise NotImplementedError('ConvTranspose2d with group > 1: the reference transposes the whole filter '
                                  '(planer/layer.py:34), which is only meaningful for group == 1')
    dt = _compute_dtype(x, K)
    xin = _as_nhwc(x, dt)
    low_h, low_w, sh_, sw_ = ops.convtranspose_geometry(xin.shape, K.shape, strides, dilations, pads, output_padding)
    c = xin.shape[1]
    cpad = _round_up(c, 16) if dt == np.float16 and c % 16 else c
    buf = B.empty((xin.shape[0], cpad, sh_, sw_), dt, 'nhwc')
    if cpad != c:                                   # padded channels must be zero for the conv
        xin = B.to_nhwc(B.to_flat(xin), dt, cpad)
    ops.zero_stuff_into(xin, buf, low_h, low_w, strides)
    stuffed = DeviceArray(buf.buf, (xin.shape[0], c, sh_, sw_), dt, 'nhwc', ld=cpad, offset=buf.offset) if cpad != c else buf
    return Conv2d(stuffed, _flipped(K), B_, 1, (1, 1), dilations, (0, 0, 0, 0))


def upsample_factors(k):
    """planer/layer.py:80-82: last two ONNX scales, truncated to int."""
    k = k.get() if isinstance(k, DeviceArray) else np.asarray(k)
    if k.size == 0:
        raise NotImplementedError('UpSample with empty scales hits an undefined name in the reference '
                                  '(planer/layer.py:81); not supported')
    f = k.reshape(-1)[-2:].astype(int).tolist()
    return int(f[0]), int(f[1])


def _int_scales(k):
    """Last two ONNX scales as integers, or None when they are not whole numbers."""
    k = k.get() if isinstance(k, DeviceArray) else np.asarray(k)
    f = k.reshape(-1)[-2:].astype(np.float64)
    return (int(f[0]), int(f[1])) if f.size == 2 and np.all(f == np.floor(f)) and np.all(f >= 1) else None


def _upsample(x, fh, fw, mode, what):
    x = _as_nhwc(x)
    n, c, h, w = x.shape
    y = B.empty((n, c, h * fh, w * fw), x.dtype, 'nhwc')
    if mode == 'nearest':
        return ops.upsample_into(x, y, fh, fw)
    if mode == 'linear' and fh >= 2 and fw >= 2:
        return ops.upsample_linear_into(x, y, fh, fw)
    raise NotImplementedError("%s: mode %r with factors (%d, %d) is not implemented (nearest, or linear with both factors >= 2)"
                              % (what, mode, fh, fw))


def UpSample(x, k, mode='nearest'):
    """planer/layer.py:80-82 -> planer/util.py:212-219: integer factors; 'nearest' (util.py:184-192, zero pixel shift) or
    'linear' (util.py:133-153, edge-replicated 4-tap blend)."""
    fh, fw = upsample_factors(k)
    return _upsample(x, fh, fw, mode, 'UpSample')


def Resize(x, roi, k, size=None, mode='nearest', coordinate_transformation_mode='half_pixel',
           nearest_mode='round_prefer_floor'):
    """planer/layer.py:84-88 for whole-number scales: the reference then takes the same two paths as UpSample
    (planer/util.py:212-219).  Nearest is implemented for the mode pairs whose pixel shift is zero (util.py:155-170: the
    ONNX defaults, asymmetric+floor, ...); fractional scales (util.py:194-210) are not."""
    if k is None or getattr(k, 'size', 0) == 0:
        raise NotImplementedError('Resize with `sizes` instead of `scales` is not implemented')
    f = _int_scales(k)
    if f is None:
        if mode != 'linear':
            raise NotImplementedError('Resize nearest with fractional scales truncates them in the reference '
                                      '(planer/util.py:213,216); not implemented')
        kk = (k.get() if isinstance(k, DeviceArray) else np.asarray(k)).reshape(-1)[-2:].tolist()
        x = _as_nhwc(x)
        n, c, h, w = x.shape
        size = resize_size((h, w), kk)
        return ops.resize_linear_into(x, B.empty((n, c) + size, x.dtype, 'nhwc'))
    if mode == 'nearest' and (nearest_shift(f[0], coordinate_transformation_mode, nearest_mode) or
                              nearest_shift(f[1], coordinate_transformation_mode, nearest_mode)):
        raise NotImplementedError('Resize nearest with a non-zero pixel shift (%s, %s) is not implemented'
                                  % (coordinate_transformation_mode, nearest_mode))
    return _upsample(x, f[0], f[1], mode, 'Resize')


def resize_size(hw, scales):
    """Output size of a fractional-scale resize (planer/util.py:214)."""
    return int(round(scales[0] * hw[0])), int(round(scales[1] * hw[1]))


def nearest_shift(k, trans_mode, round_mode):
    """Pixel shift the reference applies after a nearest upsample for an ONNX mode pair (planer/util.py:155-170): the
    position of the first source index that maps to 0."""
    idx = np.arange(-64, 64).astype(np.float64)
    if trans_mode == 'half_pixel': idx = (idx + 0.5) / k - 0.5
    if trans_mode == 'asymmetric': idx = idx / k
    idx = {'round_prefer_floor': lambda v: np.round(v - 1e-3), 'round_prefer_ceil': lambda v: np.round(v + 1e-3),
           'ceil': np.ceil, 'floor': np.floor}.get(round_mode, lambda v: v)(idx)
    return int(np.argmax(idx.astype(np.int16) == 0) - 64)


def Concatenate(*xs, axis=0):
    """planer/layer.py:90-91 for 4-D operands along the channel axis (axis=1)."""
    if axis != 1 or any(x.ndim != 4 for x in xs):
        raise NotImplementedError('Concatenate: only channel concat (axis=1) of 4-D tensors is on the B200 hot path')
    xs = [_as_nhwc(x) for x in xs]
    n, _, h, w = xs[0].shape
    ctot = sum(x.shape[1] for x in xs)
    y = B.empty((n, ctot, h, w), xs[0].dtype, 'nhwc')
    c0 = 0
    for x in xs:
        ops.copy_channels(x, ops.channel_slice(y, c0, x.shape[1]))
        c0 += x.shape[1]
    return y


def GlobalAveragePool(x):
    """planer/layer.py:77-78: mean over H, W with keepdims."""
    x = _as_nhwc(x)
    n, c = x.shape[:2]
    return ops.gap_into(x, B.empty((n, c, 1, 1), x.dtype, 'nhwc'))


def Identity(x): return x


def Return(*x):
    """planer/layer.py:260."""
    return x


class _HotPathOnly(dict):
    def __missing__(self, key):
        raise NotImplementedError("operator %r is not on the B200 hot path (SURVEY section 8) and planer_b200 "
                                  "has no CPU fallback" % (key,))


layer_map = _HotPathOnly({
    'dense': Dense, 'conv': Conv2d, 'relu': ReLU, 'leakyrelu': LeakyReLU, 'batchnorm': BatchNorm,
    'flatten': Flatten, 'sigmoid': Sigmoid, 'maxpool': Maxpool, 'upsample': UpSample,
    'concat': Concatenate, 'add': Add, 'gap': GlobalAveragePool, 'identity': Identity, 'return': Return,
    # SURVEY 8f rank 2 (the callers either side of the path): same kernels, same parity bar
    'averagepool': AveragePool, 'convtranspose': ConvTranspose2d, 'hardsigmoid': HardSigmoid, 'clip': Clip, 'softmax': Softmax,
    'resize': Resize,
})
"""Hot-path subset of planer/layer.py:262-281."""
