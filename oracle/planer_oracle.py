"""CPU oracle for Planer's per-layer forward hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithms (Image-Py/planer @ 39174495,
``/root/reference/planer/{layer,util,net}.py``).  It is *not* part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / reported CPU baseline.  The product package
``planer_b200`` never imports anything from ``oracle/``.

Parity pin: the reference ships no golden vectors and no tests, so the
oracle is pinned against outputs of the reference itself, generated in the authoring container
by ``oracle/gen_golden.py`` (imports ``/root/reference``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` replays every fixture through this file; where the fixture was
produced by the fp32 numpy path the comparison is bit-exact.

Every function cites the reference lines it follows.  The arithmetic (order of roundings,
padding values, output-size formulas, in-place aliasing) is kept identical; the code structure
is our own.
"""
import numpy as np

# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------

def out_size(n_in, pad_lo, pad_hi, k, dil, stride):
    """Output extent of a strided/dilated window walk.

    reference: planer/util.py:25-26 (conv_for) -- ``(hi + sum(pads) - (h-1)*dh - 1 + strh)//strh``;
    pooling uses the same formula with dil == 1 (planer/util.py:84-85).
    """
    return (n_in + pad_lo + pad_hi - (k - 1) * dil - 1 + stride) // stride


def zero_pad_hw(x, pads):
    """Zero padding of the two trailing axes of an NCHW array.

    reference: planer/util.py:4-10 (pad).  The reference allocates ``h + 2*top`` rows and
    ``w + 2*left`` columns, i.e. it pads *symmetrically with the leading pad* and ignores the
    trailing pad.  We restate exactly that: ``pads`` is
    ``(top, left, bottom, right)`` and only ``top``/``left`` are used for the allocation.
    """
    top, left, bottom, right = pads
    if top == bottom == left == right == 0:
        return x
    n, c, h, w = x.shape
    y = np.zeros((n, c, h + 2 * top, w + 2 * left), dtype=x.dtype)
    y[:, :, top:top + h, left:left + w] = x
    return y


# --------------------------------------------------------------------------------------
# Conv2d / Dense  (the FLOPs)
# --------------------------------------------------------------------------------------

def conv2d(x, K, B=None, group=1, strides=(1, 1), dilations=(1, 1), pads=(0, 0, 0, 0)):
    """NCHW convolution = zero-pad -> materialised im2col -> ONE matmul -> (+bias in place).

    reference: planer/layer.py:22-26 (Conv2d, numpy branch) + planer/util.py:17-44 (conv_for).
      * im2col row order is (c, r, s) with c outermost, matching ``K.reshape(Co, -1)``
        (util.py:33,41-42); columns are (n, oh, ow).
      * GEMM is ``K[Co, C*kh*kw] @ col[C*kh*kw, N*oh*ow]`` (util.py:43); for ``group > 1`` a
        batched matmul over groups (util.py:41-43).
      * result dtype is the matmul dtype (fp16 in -> fp16 out, fp32 accumulate inside numpy).
      * bias is added afterwards with ``np.add(out, B.reshape(1,-1,1,1), out=out)``
        (layer.py:26): a second rounding in fp16.
    Returned array is the same TRANSPOSED VIEW of the (Co, N, oh, ow) GEMM result the reference returns (util.py:44): the
    memory layout is part of the parity contract, because numpy's pairwise summation (GlobalAveragePool) and the BLAS path
    matmul picks (Dense after gap -> flatten) depend on the strides of what they are handed.
    """
    sh, sw = strides
    dh, dw = dilations
    co, cg, kh, kw = K.shape
    n, c, h, w = x.shape
    xp = zero_pad_hw(x, pads)
    oh = out_size(h, pads[0], pads[2], kh, dh, sh)
    ow = out_size(w, pads[1], pads[3], kw, dw, sw)
    # im2col: col[c, r*kw+s, n, oh, ow] = xp[n, c, oh*sh + r*dh, ow*sw + s*dw]   (util.py:33-38)
    col = np.zeros((c, kh * kw, n, oh, ow), dtype=x.dtype)
    xt = xp.transpose(1, 0, 2, 3)                                   # NCHW -> CNHW (util.py:30)
    for r in range(kh):
        for s in range(kw):
            col[:, r * kw + s] = xt[:, :, r * dh:r * dh + oh * sh:sh, s * dw:s * dw + ow * sw:sw]
    if group == 1:
        out = np.matmul(K.reshape(co, -1), col.reshape(c * kh * kw, -1))
    else:
        out = np.matmul(K.reshape(group, co // group, -1),
                        col.reshape(group, (c // group) * kh * kw, -1))
    out = out.reshape(co, n, oh, ow).transpose(1, 0, 2, 3)
    if B is not None:
        np.add(out, B.reshape(1, -1, 1, 1), out=out)
    return out


def convtranspose2d(x, K, B=None, strides=(2, 2), dilations=(1, 1), pads=(0, 0, 0, 0), output_padding=(0, 0), group=1):
    """Transposed convolution = zero-stuffing + ``conv2d`` with the filter transposed (in <-> out) and flipped.

    reference: planer/layer.py:28-34 (ConvTranspose2d).  K is (C_in, C_out, kh, kw); the stuffed buffer has
    ``(k-1)*d - pad`` zeros in front, ``(k-1)*d - pad + output_padding`` behind and ``stride-1`` zeros between pixels
    (layer.py:30-33); the convolution runs at stride 1 without padding (layer.py:34).
    """
    n, c, h, w = x.shape
    (s1, s2), (d1, d2), (kh, kw) = strides, dilations, K.shape[2:]
    low_h, high_h = (kh - 1) * d1 - pads[0], (kh - 1) * d1 - pads[2] + output_padding[0]
    low_w, high_w = (kw - 1) * d2 - pads[1], (kw - 1) * d2 - pads[3] + output_padding[1]
    buf = np.zeros((n, c, (h - 1) * s1 + low_h + high_h + 1, (w - 1) * s2 + low_w + high_w + 1), dtype=x.dtype)
    buf[:, :, low_h:buf.shape[2] - high_h:s1, low_w:buf.shape[3] - high_w:s2] = x
    return conv2d(buf, K.transpose(1, 0, 2, 3)[:, :, ::-1, ::-1], B, group, (1, 1), dilations)


def dense(x, K, B, shp=None):
    """``x @ K.T + B`` with K stored (out, in).  reference: planer/layer.py:15-18 (Dense)."""
    y = np.matmul(x, K.T)
    y += B.reshape((1, -1))
    return y


def matmul(x, y):
    """reference: planer/layer.py:20 (MatMul)."""
    return np.matmul(x, y)


# --------------------------------------------------------------------------------------
# elementwise companions
# --------------------------------------------------------------------------------------

def relu(x):
    """In place ``x *= (x > 0)``; returns the *same* array object (negatives become -0.0).
    reference: planer/layer.py:44-46 (ReLU, plain-numpy branch)."""
    return np.multiply(x, x > 0, out=x)


def leakyrelu(x, alpha=0.2):
    """``x * ((x>0)*(1-alpha) + alpha)`` in the dtype of x; new array.
    reference: planer/layer.py:48-51 (LeakyReLU)."""
    a, b = np.array(alpha, x.dtype), np.array(1 - alpha, x.dtype)
    y = (x > 0) * b
    y += a
    y *= x
    return y


def sigmoid(x):
    """``1 / (1 + exp(-x))`` evaluated as negate -> exp -> +1 -> reciprocal, each rounded in x.dtype.
    reference: planer/layer.py:61-64 (Sigmoid)."""
    t = -x
    np.exp(t, out=t)
    t += 1
    return np.divide(1, t, out=t)


def hardsigmoid(x, alpha=0.2, beta=0.5):
    """reference: planer/layer.py:66-69 (HardSigmoid): x*alpha, += beta, minimum 1, maximum 0 (each rounded in x's dtype)."""
    x = x * alpha
    x += beta
    x = np.minimum(x, 1, out=x)
    return np.maximum(x, 0, out=x)


def clip(x, min=0, max=1):
    """reference: planer/layer.py:247-251 (Clip, numpy branch): np.minimum(x, max) then np.maximum(x, min), in place."""
    x = np.minimum(x, max, out=x)
    return np.maximum(x, min, out=x)


def softmax(x, axis=-1):
    """reference: planer/layer.py:141-146 (Softmax, numpy branch): y = x - max; e = exp(y); y -= log(sum e); exp(y)."""
    y = x - np.max(x, axis=axis, keepdims=True)
    ey = np.exp(y)
    eX = np.sum(ey, axis=axis, keepdims=True)
    y -= np.log(eX, out=eX)
    return np.exp(y, out=y)


def add(x1, x2):
    """reference: planer/layer.py:93-95 (Add)."""
    return x1 + x2


def batchnorm(x, K, B):
    """Pre-folded batch norm ``x*K + B`` (two roundings); K,B shaped (1,C,1,1).
    reference: planer/layer.py:125-127 (BatchNorm); the fold itself is planer/io.py:76-91."""
    y = x * K
    y += B
    return y


def fold_batchnorm(gamma, beta, mean, var):
    """ONNX BatchNormalization -> per-channel scale/shift, eps hard-coded to 1e-5, host fp32.
    reference: planer/io.py:76-91 (read_onnx, BatchNormalization branch)."""
    v_inv = 1 / np.sqrt(var + 1e-5)
    shift = -gamma * mean * v_inv + beta
    scale = gamma * v_inv
    return scale.reshape(1, -1, 1, 1), shift.reshape(1, -1, 1, 1)


def flatten(x):
    """reference: planer/layer.py:59 (Flatten)."""
    return x.reshape((x.shape[0], -1))


def gap(x):
    """reference: planer/layer.py:77-78 (GlobalAveragePool)."""
    return x.mean(axis=(-2, -1), keepdims=True)


def concat(*xs, axis=0):
    """reference: planer/layer.py:90-91 (Concatenate)."""
    return np.concatenate(xs, axis=axis)


def ret(*x):
    """reference: planer/layer.py:260 (Return)."""
    return x


# --------------------------------------------------------------------------------------
# pooling / upsampling
# --------------------------------------------------------------------------------------

def _pool(x, ufunc, w, pads, strides, seed):
    """Window reduction by kh*kw strided-slice accumulations.

    reference: planer/util.py:79-92 (pool): **zero** padding (never -inf), accumulator seeded
    with ``seed`` (0 for avg, -1e4 for max), floor-mode output size.
    """
    kh, kw = w
    sh, sw = strides
    n, c, h, ww = x.shape
    xp = zero_pad_hw(x, pads)
    oh = (h + pads[0] + pads[2] - kh + sh) // sh
    ow = (ww + pads[1] + pads[3] - kw + sw) // sw
    buf = np.zeros((n, c, oh, ow), x.dtype)
    if seed != 0:
        buf[:] = seed
    for r in range(kh):
        for s in range(kw):
            ufunc(xp[:, :, r:r + oh * sh:sh, s:s + ow * sw:sw], buf, out=buf)
    return buf


def maxpool(x, w=(2, 2), pads=(0, 0, 0, 0), strides=(2, 2)):
    """reference: planer/layer.py:71-72 (Maxpool) -> planer/util.py:94-95 (maxpool, seed -1e4)."""
    return _pool(x, np.maximum, w, pads, strides, -1e4)


def avgpool(x, w=(2, 2), pads=(0, 0, 0, 0), strides=(2, 2)):
    """reference: planer/layer.py:74-75 (AveragePool) -> planer/util.py:97-100 (avgpool)."""
    y = _pool(x, np.add, w, pads, strides, 0)
    y /= w[0] * w[1]
    return y


def _nearest_shift(k, trans_mode, round_mode):
    """Pixel shift implied by an ONNX (coordinate_transformation_mode, nearest_mode) pair.
    reference: planer/util.py:155-170 (offset)."""
    idx = np.arange(-64, 64)
    if trans_mode == 'half_pixel':
        idx = (idx + 0.5) / k - 0.5
    if trans_mode == 'asymmetric':
        idx = idx / k
    if round_mode == 'round_prefer_floor':
        idx = np.round(idx - 1e-3)
    if round_mode == 'round_prefer_ceil':
        idx = np.round(idx + 1e-3)
    if round_mode == 'ceil':
        idx = np.ceil(idx)
    if round_mode == 'floor':
        idx = np.floor(idx)
    idx = idx.astype(np.int16)
    return int(np.argmax(idx == 0) - 64)


def _shift_replicate(img, dr, dc):
    """Shift an image by (dr, dc) pixels replicating the border.
    reference: planer/util.py:172-182 (pix_offset), including its sequential in-place writes."""
    n, c, h, w = img.shape
    if dr == dc == 0:
        return img
    if dr >= 0:
        r_dst, r_src, r_fill, r_from = (dr, h), (0, h - dr), (0, dr), 0
    else:
        r_dst, r_src, r_fill, r_from = (0, h + dr), (-dr, h), (h + dr, h), h - 1
    if dc >= 0:
        c_dst, c_src, c_fill, c_from = (dc, w), (0, w - dc), (0, dc), 0
    else:
        c_dst, c_src, c_fill, c_from = (0, w + dc), (-dc, w), (w + dc, w), w - 1
    img[:, :, r_dst[0]:r_dst[1], c_dst[0]:c_dst[1]] = img[:, :, r_src[0]:r_src[1], c_src[0]:c_src[1]]
    img[:, :, r_fill[0]:r_fill[1], :] = img[:, :, r_from:r_from + 1, :]
    img[:, :, :, c_fill[0]:c_fill[1]] = img[:, :, :, c_from:c_from + 1]
    return img


def upsample_nearest(x, k, trans_mode='half-pixel', round_mode='round_prefer_ceil'):
    """Integer-factor nearest upsample: ``out[..., r::kh, c::kw] = x`` for every (r, c), followed by
    the ONNX-mode pixel shift.  reference: planer/util.py:184-192 (upsample_nearest)."""
    n, c, h, w = x.shape
    out = np.zeros((n, c, h * k[0], w * k[1]), dtype=x.dtype)
    for r in range(k[0]):
        for s in range(k[1]):
            out[:, :, r::k[0], s::k[1]] = x
    return _shift_replicate(out, _nearest_shift(k[0], trans_mode, round_mode),
                            _nearest_shift(k[1], trans_mode, round_mode))


def _bilinear_taps(k):
    """(4, kh*kw) blend weights of the four corners (left-top, right-top, left-bottom, right-bottom) for every output position
    inside one input cell -- computed in FLOAT16 like the reference.  reference: planer/util.py:121-131 (make_upmat);
    kh == 1 or kw == 1 degenerate to two taps there and are not restated (the B200 path requires both factors >= 2)."""
    ys = np.linspace(0.5 / k[0], 1 - 0.5 / k[0], k[0], dtype=np.float16)[:, None]
    xs = np.linspace(0.5 / k[1], 1 - 0.5 / k[1], k[1], dtype=np.float16)[None, :]
    return np.vstack([((1 - xs) * (1 - ys)).reshape(1, -1), (xs * (1 - ys)).reshape(1, -1),
                      ((1 - xs) * ys).reshape(1, -1), (xs * ys).reshape(1, -1)])


def upsample_bilinear(x, k):
    """Integer-factor bilinear upsample: replicate the border by one pixel, blend every 2x2 neighbourhood into a kh x kw
    block with ONE matmul (cells, 4) @ (4, kh*kw), interleave the blocks and crop kh//2, kw//2 on each side.
    reference: planer/util.py:133-153 (upsample_blinear), both factors >= 2."""
    n, c, h, w = x.shape
    assert k[0] >= 2 and k[1] >= 2
    xp = np.concatenate((x[:, :, :1, :], x, x[:, :, -1:, :]), axis=2)
    xp = np.concatenate((xp[:, :, :, :1], xp, xp[:, :, :, -1:]), axis=3)
    corners = [xp[:, :, :-1, :-1], xp[:, :, :-1, 1:], xp[:, :, 1:, :-1], xp[:, :, 1:, 1:]]
    cells = np.concatenate([t[:, :, :, :, None] for t in corners], axis=-1)
    out = np.matmul(cells.reshape((-1, 4)), _bilinear_taps(k))
    hh, ww = h + 1, w + 1
    out = out.reshape((-1, ww, k[0], k[1])).transpose((0, 2, 1, 3)).reshape((n, c, hh * k[0], ww * k[1]))
    return out[:, :, k[0] // 2:h * k[0] + k[0] // 2, k[1] // 2:w * k[1] + k[1] // 2]


def upsample_size(x, size):
    """Bilinear resize to an arbitrary (H, W): centre-aligned source coordinates computed IN THE IMAGE DTYPE, clipped to the
    image, gather + lerp along columns, then along rows.  reference: planer/util.py:194-210 (upsample_size)."""
    lead, (h, w) = x.shape[:-2], x.shape[-2:]
    kh, kw = size[0] / h, size[1] / w
    rs = np.linspace(-0.5 + 0.5 / kh, h - 0.5 - 0.5 / kh, size[0], dtype=x.dtype)
    cs = np.linspace(-0.5 + 0.5 / kw, w - 0.5 - 0.5 / kw, size[1], dtype=x.dtype)
    rs = np.clip(rs, 0, h - 1, out=rs)
    cs = np.clip(cs, 0, w - 1, out=cs)
    ra = np.floor(np.clip(rs, 0, h - 1.001)).astype(int)
    ca = np.floor(np.clip(cs, 0, w - 1.001)).astype(int)
    rs -= ra
    cs -= ca
    rs = rs.reshape(-1, 1)
    img = x.reshape(-1, h, w)
    cols = img[:, :, ca] * (1 - cs) + img[:, :, ca + 1] * cs
    out = cols[:, ra, :] * (1 - rs) + cols[:, ra + 1, :] * rs
    return out.reshape(lead + tuple(size))


def _upsample(x, k, mode, trans_mode, round_mode):
    """reference: planer/util.py:212-219 (upsample) for whole-number factors."""
    kint = [int(k[0]), int(k[1])]
    if mode == 'nearest':
        return upsample_nearest(x, kint, trans_mode, round_mode)
    if mode == 'linear' and k[0] == int(k[0]) and k[1] == int(k[1]):
        return upsample_bilinear(x, kint)
    if mode == 'linear':
        return upsample_size(x, (int(round(k[0] * x.shape[2])), int(round(k[1] * x.shape[3]))))
    raise NotImplementedError('oracle: unknown resize mode %r' % mode)


def upsample(x, k, mode='nearest'):
    """``k`` is the ONNX scales tensor; its last two entries, truncated to int, are the factors.
    reference: planer/layer.py:80-82 (UpSample) -> planer/util.py:212-216 (upsample; note the
    default mode strings 'half-pixcel' / 'round_prefer_ceil' select a zero shift)."""
    kk = np.asarray(k)[-2:].astype(int).tolist()
    return _upsample(x, kk, mode, 'half-pixcel', 'round_prefer_ceil')


def resize(x, roi, k, size=None, mode='nearest', coordinate_transformation_mode='half_pixel',
           nearest_mode='round_prefer_floor'):
    """reference: planer/layer.py:84-88 (Resize) with a scales tensor."""
    return _upsample(x, np.asarray(k)[-2:].tolist(), mode, coordinate_transformation_mode, nearest_mode)


# --------------------------------------------------------------------------------------
# operator table + graph interpreter
# --------------------------------------------------------------------------------------

layer_map = {
    'conv': conv2d, 'dense': dense, 'matmul': matmul, 'relu': relu, 'leakyrelu': leakyrelu,
    'sigmoid': sigmoid, 'add': add, 'batchnorm': batchnorm, 'flatten': flatten, 'gap': gap,
    'concat': concat, 'maxpool': maxpool, 'averagepool': avgpool, 'upsample': upsample,
    'convtranspose': convtranspose2d, 'hardsigmoid': hardsigmoid, 'clip': clip, 'softmax': softmax, 'resize': resize, 'return': ret,
}
"""Hot-path subset of planer/layer.py:262-281 (layer_map)."""


class OracleNet:
    """Sequential interpreter over Planer's JSON IR.

    reference: planer/net.py:5-101 (Net).  ``load_json`` (net.py:10-24) builds the operator
    closures and the liveness table, ``load_weights`` (net.py:83-88) slices the flat uint8
    blob, ``half`` (net.py:26-29) casts every fp32 init to fp16, ``forward`` (net.py:37-72)
    walks ``flow`` -- in a chained flow only the first layer reads the listed inputs, later
    layers read the previous output (net.py:46-50) -- and drops dead values (net.py:51-53).
    """

    def __init__(self, table=None):
        self.table = dict(layer_map if table is None else table)
        self.weights, self.body, self.flow, self.life, self.timer = [], [], [], {}, {}

    def load_json(self, inputs, inits, body, flow):
        self.body = [(name, (kind, dict(para))) for name, kind, para in body]
        self.life = {}
        for i, (xs, _, _) in enumerate(flow):
            for k in ([xs] if isinstance(xs, str) else xs):
                self.life[k] = i
        self.weights = [np.zeros(shape, dtype=dt) for _, shape, dt in inits]
        self.input, self.inits = inputs, [i[0] for i in inits]
        self.layer, self.flow = body, flow

    def load_weights(self, blob):
        pos, blob = 0, np.asarray(blob).view(np.uint8)
        for w in self.weights:
            raw = w.reshape(-1).view(np.uint8)
            raw[:] = blob[pos:pos + raw.size]
            pos += raw.size

    def half(self):
        self.weights = [w.astype('float16') if w.dtype == np.float32 else w for w in self.weights]

    def forward(self, *x, trace=None):
        import time
        ops = dict(self.body)
        val = {'None': None}
        val.update(zip(self.inits, self.weights))
        val.update(zip(self.input, x))
        y = None
        for i, (xs, names, y) in enumerate(self.flow):
            names = names if isinstance(names, list) else [names]
            for j, name in enumerate(names):
                src = xs if j == 0 else y
                args = [val[src]] if isinstance(src, str) else [val.get(k) for k in src]
                for k in set(xs if isinstance(xs, list) else [xs]):
                    if k in val and self.life[k] <= i:
                        del val[k]
                kind, para = ops[name]
                t0 = time.time()
                out = self.table[kind](*args, **para)
                self.timer[kind] = self.timer.get(kind, 0) + time.time() - t0
                if isinstance(y, str):
                    val[y] = out
                else:
                    val.update(zip(y, out))
                if trace is not None:
                    trace.append((name, kind, out))
        return val[y]

    def __call__(self, *x):
        if isinstance(x[0], dict):
            x = [x[0][k] for k in self.input]
        out = self.forward(*x)
        return out[0] if isinstance(out, tuple) and len(out) == 1 else out


def build_net(model, blob, half=False):
    """model = {'input','inits','layers','flow'} dict + uint8 blob -> OracleNet.
    reference: planer/io.py:19-24,32-33 (read_net, json+npy branch)."""
    net = OracleNet()
    net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
    net.load_weights(blob)
    if half:
        net.half()
    return net
