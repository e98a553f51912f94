"""Destination-passing wrappers over the C ABI (one Python call = one kernel launch).

These are the primitives both the eager operator table (``layer.py``) and the fused plan (``plan.py``)
are made of.  Every function takes pre-allocated DeviceArrays, mirrors the argument meaning of the
reference function it replaces, and raises ``PlanerB200Error`` on any failure (no fallback).
"""
import ctypes as C

import numpy as np

from . import _capi
from . import backend as B
from ._capi import (ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ALGO_AUTO, ALGO_TCGEN05, ALGO_DIRECT,
                    EW_RELU, EW_LEAKY, EW_SIGMOID, EW_ADD, EW_SCALE_SHIFT, EW_CLIP, EW_HARDSIGMOID)


def out_size(n_in, pad_lo, pad_hi, k, dil, stride):
    """planer/util.py:25-26."""
    return (n_in + pad_lo + pad_hi - (k - 1) * dil - 1 + stride) // stride


def conv_out_shape(xshape, kshape, strides, dilations, pads):
    n, c, h, w = xshape
    co, _, kh, kw = kshape
    return (n, co, out_size(h, pads[0], pads[2], kh, dilations[0], strides[0]),
            out_size(w, pads[1], pads[3], kw, dilations[1], strides[1]))


def pool_out_shape(xshape, w, pads, strides):
    """planer/util.py:84-85."""
    n, c, h, ww = xshape
    return (n, c, (h + pads[0] + pads[2] - w[0] + strides[0]) // strides[0],
            (ww + pads[1] + pads[3] - w[1] + strides[1]) // strides[1])


def _epilogue(scale, shift, residual, act, alpha, res_after_act=False, out_nchw=False, out_f32=False, acc_scale=1.0,
              acc_scale_dev=None, pool_sum=None):
    keep = []
    ep = _capi.Epilogue()
    ep.out_nchw = int(bool(out_nchw))
    ep.pool_sum = pool_sum.ptr if pool_sum is not None else None
    ep.out_f32, ep.acc_scale, ep.acc_scale_dev = int(bool(out_f32)), float(acc_scale), acc_scale_dev
    ep.scale = scale.ptr if scale is not None else None
    ep.shift = shift.ptr if shift is not None else None
    if residual is not None:
        t = residual.tensor()
        keep.append(t)
        ep.residual = C.pointer(t)
    ep.act, ep.alpha, ep.res_after_act = int(act), float(alpha), int(bool(res_after_act))
    return ep, keep


def pack_weight(K, cin_pad, dtype, co_pad=None):
    """OIHW -> [Cout][kh][kw][cin_pad] in ``dtype`` (one-off per layer); ``co_pad`` > Cout appends zero filters."""
    co, cg, kh, kw = K.shape
    out = B.empty((co, kh, kw, cin_pad), dtype) if not co_pad or co_pad == co else B.zeros((co_pad, kh, kw, cin_pad), dtype)
    _capi.check(B.lib().plnr_pack_conv_weight(B.ctx(), K.ptr, _capi.dtype_code(K.dtype), out.ptr,
                                              _capi.dtype_code(dtype), co, cg, kh, kw, cin_pad),
                'plnr_pack_conv_weight')
    return out


# ---- float32 convolutions on the fp16 tensor pipe (csrc/split_f32.cu) ---------------------------------------------------
def split_conv_enabled():
    """PLNR_F32_TENSOR=0 keeps float32 convolutions on the CUDA-core FFMA kernel (conv_direct.cu)."""
    import os
    return os.environ.get('PLNR_F32_TENSOR', '1') != '0'


def split_channels(c):
    """Stored channels of the split operand: [hi C | hi C | lo C] rounded up to the tensor-core kernel's multiple of 16."""
    return (3 * c + 15) // 16 * 16


class SplitWeight:
    """A float32 filter split into fp16 (hi, lo) pairs for the tensor-core path: ``packed`` is [Cout][kh][kw][cs] fp16 with
    [hi | lo | hi] per tap, ``prescale`` the power of two the filter was multiplied by.  Owns the fp16 staging tensors of
    its input (one per input shape, zeroed once: the pad channels are never written) and the two device floats of the
    per-call activation pre-scale."""

    def __init__(self, packed, prescale, cin):
        self.packed, self.prescale, self.cin = packed, float(prescale), int(cin)
        self.cs = packed.shape[3]
        self._stage = {}

    def stage(self, x):
        key = (x.shape[0], x.shape[2], x.shape[3])
        if key not in self._stage:
            if len(self._stage) >= 4:
                self._stage.pop(next(iter(self._stage)))
            xs = B.empty((x.shape[0], self.cs, x.shape[2], x.shape[3]), np.float16, 'nhwc')
            _capi.check(B.lib().plnr_memset(B.ctx(), xs.ptr, 0, max(2 * x.shape[0] * x.shape[2] * x.shape[3] * self.cs, 1)), 'plnr_memset')
            self._stage[key] = (xs, B.zeros((2,), np.float32))
        return self._stage[key]


def pack_weight_split(K):
    """OIHW float32 filter -> (packed fp16 [Cout][kh][kw][cs], meta fp32 [prescale]) for ``SplitWeight``.  The pre-scale
    puts max|w| into [2^13, 2^14) so that the low parts stay out of fp16's subnormal range."""
    co, ci, kh, kw = K.shape
    cs = split_channels(ci)
    amax = B.empty((1,), np.float32)
    _capi.check(B.lib().plnr_absmax_f32(B.ctx(), K.ptr, int(np.prod(K.shape)), amax.ptr), 'plnr_absmax_f32')
    m = float(amax.get()[0])
    e = 0 if not np.isfinite(m) or m <= 0 else 13 - int(np.floor(np.log2(m)))
    prescale = 2.0 ** max(-100, min(100, e))
    out = B.zeros((co, kh, kw, cs), np.float16)
    _capi.check(B.lib().plnr_pack_conv_weight_split(B.ctx(), K.ptr, out.ptr, co, ci, kh, kw, cs, prescale),
                'plnr_pack_conv_weight_split')
    return out, B.asarray(np.array([prescale], np.float32))


def split_weight(packed, meta, cin):
    return SplitWeight(packed, float(meta.get()[0]), cin)


def _conv2d_split_into(x, sw, y, kh, kw, strides, dilations, pads, scale, shift, residual, act, alpha, res_after_act):
    """float32 x, y (and residual); fp16 split operands; three launches (max|x|, split, tensor-core conv with the fp32
    epilogue)."""
    xs, dyn = sw.stage(x)
    tx, txs = x.tensor(), xs.tensor()
    _capi.check(B.lib().plnr_split_f32(B.ctx(), C.byref(tx), C.byref(txs), 1.0, dyn.ptr), 'plnr_split_f32')
    d = _capi.ConvDesc(_capi.dtype_code(np.float16), kh, kw, pads[0], pads[1], pads[2], pads[3],
                       strides[0], strides[1], dilations[0], dilations[1], 1, ALGO_AUTO)
    ty = y.tensor()
    ep, keep = _epilogue(scale, shift, residual, act, alpha, res_after_act, out_f32=True,
                         acc_scale=1.0 / sw.prescale, acc_scale_dev=dyn.ptr + 4)
    _capi.check(B.lib().plnr_conv2d_fwd(B.ctx(), C.byref(d), C.byref(txs), sw.packed.ptr, C.byref(ty), C.byref(ep)),
                'plnr_conv2d_fwd (float32 on the tensor pipe)')
    return y


def pad_vector(v, n):
    """fp32 per-channel vector -> length n, zero-filled tail (channel-padded conv outputs)."""
    out = B.zeros((n,), v.dtype)
    out[0:v.shape[0]] = v
    return out


def fold_affine(bias, bn_k, bn_b, c):
    """(bias, folded-BN K, B) -> fp32 per-channel (scale, shift) of the fused epilogue."""
    dts = {a.dtype for a in (bias, bn_k, bn_b) if a is not None}
    if len(dts) > 1:   # mixed precision inits: compute the fold in fp32
        bias, bn_k, bn_b = [None if a is None else a.astype(np.float32) for a in (bias, bn_k, bn_b)]
        dts = {np.dtype(np.float32)}
    dt = dts.pop() if dts else np.dtype(np.float32)
    scale, shift = B.empty((c,), np.float32), B.empty((c,), np.float32)
    p = lambda a: a.ptr if a is not None else None
    _capi.check(B.lib().plnr_fold_affine(B.ctx(), p(bias), p(bn_k), p(bn_b), _capi.dtype_code(dt), scale.ptr,
                                         shift.ptr, c), 'plnr_fold_affine')
    return scale, shift


def nchw_out_tensor(y):
    """plnr_tensor header of a FLAT (NCHW) 4-D array used as a conv output with ``out_nchw`` (extents only; ld = c)."""
    n, c, h, w = y.shape
    return _capi.Tensor(y.ptr, n, h, w, c, c, 0)


def conv2d_out_nchw_supported(x, yshape, kh, kw, strides, dilations, pads):
    """True when the conv epilogue can write the dense NCHW result itself (graph exit without a transpose launch)."""
    d = _capi.ConvDesc(_capi.dtype_code(x.dtype), kh, kw, pads[0], pads[1], pads[2], pads[3],
                       strides[0], strides[1], dilations[0], dilations[1], 1, ALGO_AUTO)
    tx = x.tensor()
    ty = _capi.Tensor(256, yshape[0], yshape[2], yshape[3], yshape[1], yshape[1], 0)
    return bool(B.lib().plnr_conv2d_out_nchw_supported(C.byref(d), C.byref(tx), C.byref(ty)))


def conv2d_pool_parts(x, yshape, kh, kw, strides, dilations, pads):
    """32-position parts per image when GlobalAveragePool can be folded into this convolution's epilogue
    (plnr_epilogue.pool_sum), 0 when it cannot."""
    d = _capi.ConvDesc(_capi.dtype_code(x.dtype), kh, kw, pads[0], pads[1], pads[2], pads[3],
                       strides[0], strides[1], dilations[0], dilations[1], 1, ALGO_AUTO)
    tx = x.tensor()
    ty = _capi.Tensor(256, yshape[0], yshape[2], yshape[3], yshape[1], yshape[1], 0)
    return int(B.lib().plnr_conv2d_pool_parts(C.byref(d), C.byref(tx), C.byref(ty)))


def conv2d_into(x, w_packed, y, kh, kw, strides, dilations, pads, groups=1, scale=None, shift=None,
                residual=None, act=ACT_NONE, alpha=0.0, algo=ALGO_AUTO, res_after_act=False, out_nchw=False, pool_sum=None):
    """``out_nchw``: y is a flat NCHW array and the epilogue writes it directly (plnr_epilogue.out_nchw).  ``pool_sum``: a
    float32 (n, parts, cout) array that receives the partial sums of the GlobalAveragePool consuming this layer INSTEAD of y
    (plnr_epilogue.pool_sum; parts from ``conv2d_pool_parts``).  ``w_packed`` is the array from ``pack_weight`` or, for
    float32 on the tensor pipe, a ``SplitWeight``."""
    if isinstance(w_packed, SplitWeight):
        assert groups == 1 and not out_nchw and pool_sum is None and x.dtype == np.float32 and y.dtype == np.float32
        return _conv2d_split_into(x, w_packed, y, kh, kw, strides, dilations, pads, scale, shift, residual, act, alpha,
                                  res_after_act)
    d = _capi.ConvDesc(_capi.dtype_code(x.dtype), kh, kw, pads[0], pads[1], pads[2], pads[3],
                       strides[0], strides[1], dilations[0], dilations[1], groups, algo)
    tx, ty = x.tensor(), (nchw_out_tensor(y) if out_nchw else y.tensor())
    ep, keep = _epilogue(scale, shift, residual, act, alpha, res_after_act, out_nchw, pool_sum=pool_sum)
    rc = B.lib().plnr_conv2d_fwd(B.ctx(), C.byref(d), C.byref(tx), w_packed.ptr, C.byref(ty), C.byref(ep))
    if rc != 0:
        desc = lambda a: None if a is None else '%s %s ld=%s coff=%s' % (a.shape, a.layout, a.ld, a.coff)
        _capi.check(rc, 'plnr_conv2d_fwd [x %s | y %s | residual %s | k %dx%d s %s]' % (desc(x), desc(y), desc(residual), kh, kw,
                                                                                     tuple(strides)))
    return y


def conv2d_shortcut_supported(dtype, xshape, x2shape, stride2, yshape, kh, kw, strides, dilations, pads):
    """True when conv(x) + conv1x1_stride2(x2) can run as ONE launch (plnr_conv2d_shortcut_fwd); shapes are logical
    (N, C, H, W) of dense pixel-major tensors."""
    d = _capi.ConvDesc(_capi.dtype_code(dtype), kh, kw, pads[0], pads[1], pads[2], pads[3],
                       strides[0], strides[1], dilations[0], dilations[1], 1, ALGO_AUTO)
    mk = lambda shp: _capi.Tensor(256, shp[0], shp[2], shp[3], shp[1], shp[1], 0)     # dummy aligned pointer
    tx, tx2, ty = mk(xshape), mk(x2shape), mk(yshape)
    return bool(B.lib().plnr_conv2d_shortcut_supported(C.byref(d), C.byref(tx), C.byref(tx2), stride2, C.byref(ty)))


def conv2d_shortcut_into(x, w_cat, x2, stride2, y, kh, kw, strides, dilations, pads, scale=None, shift=None,
                         act=ACT_NONE, alpha=0.0):
    """y = act((conv(x, W) + conv1x1(x2, W2, stride2)) * scale + shift); w_cat = packed W with W2 appended along K."""
    d = _capi.ConvDesc(_capi.dtype_code(x.dtype), kh, kw, pads[0], pads[1], pads[2], pads[3],
                       strides[0], strides[1], dilations[0], dilations[1], 1, ALGO_AUTO)
    tx, tx2, ty = x.tensor(), x2.tensor(), y.tensor()
    ep, keep = _epilogue(scale, shift, None, act, alpha, False)
    _capi.check(B.lib().plnr_conv2d_shortcut_fwd(B.ctx(), C.byref(d), C.byref(tx), w_cat.ptr, C.byref(tx2), stride2,
                                                 C.byref(ty), C.byref(ep)), 'plnr_conv2d_shortcut_fwd')
    return y


def conv2d_algo(x, y, kh, kw, strides, dilations, pads, groups=1):
    d = _capi.ConvDesc(_capi.dtype_code(x.dtype), kh, kw, pads[0], pads[1], pads[2], pads[3],
                       strides[0], strides[1], dilations[0], dilations[1], groups, ALGO_AUTO)
    tx, ty = x.tensor(), y.tensor()
    return B.lib().plnr_conv2d_algo(C.byref(d), C.byref(tx), C.byref(ty))


def dense_into(x, w, y, scale=None, shift=None, residual=None, act=ACT_NONE, alpha=0.0, algo=ALGO_AUTO,
               res_after_act=False):
    m, k = x.shape
    n = w.shape[0]
    ep, keep = _epilogue(scale, shift, residual, act, alpha, res_after_act)
    _capi.check(B.lib().plnr_dense_fwd(B.ctx(), _capi.dtype_code(x.dtype), x.ptr, w.ptr, y.ptr, m, n, k,
                                       C.byref(ep), algo), 'plnr_dense_fwd')
    return y


def maxpool_into(x, y, w, pads, strides):
    tx, ty = x.tensor(), y.tensor()
    _capi.check(B.lib().plnr_maxpool2d(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), C.byref(ty), w[0], w[1],
                                       pads[0], pads[1], strides[0], strides[1]), 'plnr_maxpool2d')
    return y


def avgpool_into(x, y, w, pads, strides):
    tx, ty = x.tensor(), y.tensor()
    _capi.check(B.lib().plnr_avgpool2d(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), C.byref(ty), w[0], w[1],
                                       pads[0], pads[1], strides[0], strides[1]), 'plnr_avgpool2d')
    return y


def convtranspose_geometry(x_shape, k_shape, strides, dilations, pads, output_padding):
    """planer/layer.py:29-32: (low_h, low_w, stuffed H, stuffed W) of the zero-stuffed buffer the stride-1 conv then reads."""
    n, c, h, w = x_shape
    kh, kw = k_shape[2:]
    low_h, high_h = (kh - 1) * dilations[0] - pads[0], (kh - 1) * dilations[0] - pads[2] + output_padding[0]
    low_w, high_w = (kw - 1) * dilations[1] - pads[1], (kw - 1) * dilations[1] - pads[3] + output_padding[1]
    if min(low_h, high_h, low_w, high_w) < 0:
        raise NotImplementedError('ConvTranspose2d: pads larger than (k-1)*dilation crop the stuffed input (the reference\'s '
                                  'negative slice bounds, planer/layer.py:32-33); not supported')
    return low_h, low_w, (h - 1) * strides[0] + low_h + high_h + 1, (w - 1) * strides[1] + low_w + high_w + 1


def zero_stuff_into(x, y, low_h, low_w, strides):
    tx, ty = x.tensor(), y.tensor()
    _capi.check(B.lib().plnr_zero_stuff(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), C.byref(ty), low_h, low_w,
                                        strides[0], strides[1]), 'plnr_zero_stuff')
    return y


def flip_weight(K):
    """K (ci, co, kh, kw) -> K.transpose(1,0,2,3)[:, :, ::-1, ::-1] as a new flat device array (planer/layer.py:34)."""
    ci, co, kh, kw = K.shape
    src = B.to_flat(K)
    out = B.empty((co, ci, kh, kw), K.dtype)
    _capi.check(B.lib().plnr_flip_weight(B.ctx(), _capi.dtype_code(K.dtype), src.ptr, out.ptr, ci, co, kh, kw), 'plnr_flip_weight')
    return out


_upmat_cache = {}


def upsample_linear_weights(fh, fw):
    """The reference's 4-tap weight table for integer-factor bilinear upsampling (planer/util.py:121-131, make_upmat):
    sample positions and products in FLOAT16, rows = (left-top, right-top, left-bottom, right-bottom); device fp32 (4, fh*fw)."""
    if (fh, fw) not in _upmat_cache:
        ys = np.linspace(0.5 / fh, 1 - 0.5 / fh, fh, dtype=np.float16)[:, None]
        xs = np.linspace(0.5 / fw, 1 - 0.5 / fw, fw, dtype=np.float16)[None, :]
        taps = [(1 - xs) * (1 - ys), xs * (1 - ys), (1 - xs) * ys, xs * ys]
        _upmat_cache[(fh, fw)] = B.asarray(np.stack([t.reshape(-1) for t in taps]).astype(np.float32))
    return _upmat_cache[(fh, fw)]


def upsample_linear_into(x, y, fh, fw, wm=None):
    """``wm``: the weight table from ``upsample_linear_weights`` when the caller owns it (an executor whose CUDA graph
    holds its address); None = the module cache (eager calls)."""
    tx, ty = x.tensor(), y.tensor()
    wm = upsample_linear_weights(fh, fw) if wm is None else wm
    _capi.check(B.lib().plnr_upsample_linear(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), C.byref(ty), fh, fw, wm.ptr),
                'plnr_upsample_linear')
    return y


def resize_linear_tables(n_in, n_out, dtype):
    """Source index / weights of every output sample along one axis, computed in the IMAGE dtype exactly like the reference
    (planer/util.py:196-206): centre-aligned coordinates, clipped; -> (lo int32, weight of lo+1, weight of lo) as fp32."""
    dtype = np.dtype(dtype)
    k = n_out / n_in
    pos = np.linspace(-0.5 + 0.5 / k, n_in - 0.5 - 0.5 / k, n_out, dtype=dtype)
    pos = np.clip(pos, 0, n_in - 1, out=pos)
    lo = np.floor(np.clip(pos, 0, n_in - 1.001)).astype(int)
    pos -= lo
    return lo.astype(np.int32), pos.astype(np.float32), (1 - pos).astype(np.float32)


_resize_cache = {}


def resize_linear_device_tables(x, y):
    """The six device tables of one (input size, output size, dtype) resize."""
    return [B.asarray(t) for t in resize_linear_tables(x.shape[2], y.shape[2], x.dtype) +
            resize_linear_tables(x.shape[3], y.shape[3], x.dtype)]


def resize_linear_into(x, y, tables=None):
    """Bilinear resize of x (n, c, h, w) to y's spatial size (planer/util.py:194-210).  ``tables``: device tables owned by
    the caller (an executor: its captured CUDA graph keeps reading them); None = a small module-level cache for eager calls
    (entries are evicted one at a time, oldest first, never while a captured graph can refer to them)."""
    if tables is None:
        key = (x.shape[2], x.shape[3], y.shape[2], y.shape[3], str(x.dtype))
        if key not in _resize_cache:
            while len(_resize_cache) >= 256:
                _resize_cache.pop(next(iter(_resize_cache)))
            _resize_cache[key] = resize_linear_device_tables(x, y)
        tables = _resize_cache[key]
    rl, rw, rw1, cl, cw, cw1 = tables
    tx, ty = x.tensor(), y.tensor()
    _capi.check(B.lib().plnr_resize_linear(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), C.byref(ty), rl.ptr, rw.ptr, rw1.ptr,
                                           cl.ptr, cw.ptr, cw1.ptr), 'plnr_resize_linear')
    return y


def upsample_into(x, y, fh, fw):
    tx, ty = x.tensor(), y.tensor()
    _capi.check(B.lib().plnr_upsample_nearest(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), C.byref(ty), fh, fw),
                'plnr_upsample_nearest')
    return y


def copy_channels(x, y):
    tx, ty = x.tensor(), y.tensor()
    _capi.check(B.lib().plnr_copy_channels(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), C.byref(ty)),
                'plnr_copy_channels')
    return y


def channel_slice(a, c0, c):
    """View of channels [c0, c0+c) of an nhwc array (zero-copy concat target)."""
    n, _, h, w = a.shape
    return B.DeviceArray(a.buf, (n, c, h, w), a.dtype, 'nhwc', ld=a.ld, coff=a.coff + c0, offset=a.offset)


def eltwise(op, x, y, p0=None, p1=None, alpha=0.0):
    """x, y: dense nhwc (ld == C) or flat arrays of equal size; per-channel params index the last (C) axis."""
    if x.layout == 'nhwc':
        assert x.ld == x.shape[1] and x.coff == 0, 'eltwise needs dense rows'
        c = x.shape[1]
    else:
        c = x.shape[-1] if x.ndim else 1
    npix = x.size // max(c, 1)
    p = lambda a: a.ptr if a is not None else None
    _capi.check(B.lib().plnr_eltwise(B.ctx(), op, _capi.dtype_code(x.dtype), x.ptr, p(p0), p(p1), y.ptr, npix, c,
                                     float(alpha)), 'plnr_eltwise')
    return y


def unary2(op, x, y, a, b):
    """CLIP (a = min, b = max) / HARDSIGMOID (a = alpha, b = beta) on dense arrays of equal size."""
    _capi.check(B.lib().plnr_unary2(B.ctx(), op, _capi.dtype_code(x.dtype), x.ptr, y.ptr, x.size, float(a), float(b)), 'plnr_unary2')
    return y


def softmax_into(x, y):
    """Softmax over the innermost stored axis: channels of a dense nhwc array, or the last axis of a flat array."""
    if x.layout == 'nhwc':
        assert x.ld == x.shape[1] and x.coff == 0, 'softmax needs dense rows'
        c = x.shape[1]
    else:
        c = x.shape[-1]
    _capi.check(B.lib().plnr_softmax(B.ctx(), _capi.dtype_code(x.dtype), x.ptr, y.ptr, x.size // c, c), 'plnr_softmax')
    return y


def gap_into(x, y):
    tx = x.tensor()
    _capi.check(B.lib().plnr_global_avgpool(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), y.ptr),
                'plnr_global_avgpool')
    return y


def gap_dense_into(x, w, y, scale=None, shift=None, act=ACT_NONE, alpha=0.0):
    """GlobalAveragePool -> Flatten -> Dense in one launch: x nhwc (n, c, h, w), w (out, c), y flat (n, out)."""
    tx = x.tensor()
    p = lambda a: a.ptr if a is not None else None
    _capi.check(B.lib().plnr_gap_dense_fwd(B.ctx(), _capi.dtype_code(x.dtype), C.byref(tx), w.ptr, p(scale), p(shift), y.ptr,
                                           w.shape[0], act, float(alpha)), 'plnr_gap_dense_fwd')
    return y


def pooled_dense_into(pool, hw, w, y, scale=None, shift=None, act=ACT_NONE, alpha=0.0):
    """The gap -> flatten -> dense tail after a convolution that pooled in its epilogue: pool float32 (n, parts, c) partial
    sums, hw pooled positions per image, w (out, c) fp16, y flat (n, out) fp16."""
    p = lambda a: a.ptr if a is not None else None
    n, parts, c = pool.shape
    _capi.check(B.lib().plnr_pooled_dense_fwd(B.ctx(), pool.ptr, n, parts, c, int(hw), w.ptr, p(scale), p(shift), y.ptr,
                                              w.shape[0], act, float(alpha)), 'plnr_pooled_dense_fwd')
    return y


def nchw_to_nhwc_into(x_flat, y, c_src=None):
    t = y.tensor()
    _capi.check(B.lib().plnr_nchw_to_nhwc(B.ctx(), x_flat.ptr, _capi.src_dtype_code(x_flat.dtype),
                                          x_flat.shape[1] if c_src is None else c_src, C.byref(t),
                                          _capi.dtype_code(y.dtype)), 'plnr_nchw_to_nhwc')
    return y


def stem_geometry(h, w, kh, kw, stride, pads):
    """Geometry of the packed first layer: a kh x kw / stride-s conv of an (h, w) image == a (T x 1) / 1 conv over
    a packed tensor of H2 rows x OW columns.  Vertical tap r of the original filter satisfies
    r - pad_top = s*e + ph  (e = packed row offset, ph = stride phase)."""
    pt, pl, pb, pr = pads
    oh, ow = out_size(h, pt, pb, kh, 1, stride), out_size(w, pl, pr, kw, 1, stride)
    e_min, e_max = (-pt) // stride, (kh - 1 - pt) // stride
    T = e_max - e_min + 1
    h2 = (h + stride - 1) // stride
    pad_t2 = -e_min
    pad_b2 = oh - h2 - pad_t2 + T - 1
    return dict(oh=oh, ow=ow, T=T, h2=h2, pad_t2=pad_t2, pad_b2=pad_b2, e_min=e_min)


def stem_pack_weight(K, stride, pads, cp):
    """Host-side (tiny, load-time) re-ordering of an OIHW stem filter into the packed [Cout][T][1][cp] layout that
    matches ``plnr_stem_pack``:  W2[co, e, (ph*kw + sx)*C + c] = K[co, c, s*(e+e_min) + ph + pad_top, sx]."""
    co, c, kh, kw = K.shape
    pt = pads[0]
    e_min, e_max = (-pt) // stride, (kh - 1 - pt) // stride
    T = e_max - e_min + 1
    out = np.zeros((co, T, 1, cp), K.dtype)
    for e in range(T):
        for ph in range(stride):
            r = stride * (e + e_min) + ph + pt
            if 0 <= r < kh:
                for sx in range(kw):
                    base = (ph * kw + sx) * c
                    out[:, e, 0, base:base + c] = K[:, :, r, sx]
    return out


def stem_pack_into(x_flat, y, kw, stride, pad_l):
    n, c, h, w = x_flat.shape
    t = y.tensor()
    _capi.check(B.lib().plnr_stem_pack(B.ctx(), x_flat.ptr, _capi.src_dtype_code(x_flat.dtype), n, c, h, w, C.byref(t), kw,
                                       stride, pad_l), 'plnr_stem_pack')
    return y


def stem_pool_supported(dtype, c, h, w, cout, kh, kw, stride, pads, act, pool_w, pool_strides, pool_pads):
    """True when the fused first-layer kernel (conv + scale/shift + ReLU + 3x3/s2/p1 maxpool) applies."""
    if tuple(pool_w) != (3, 3) or tuple(pool_strides) != (2, 2) or tuple(pool_pads) != (1, 1, 1, 1):
        return False
    return bool(B.lib().plnr_stem_pool_supported(_capi.dtype_code(dtype), c, h, w, cout, kh, kw, stride, pads[0], pads[1],
                                                 pads[2], pads[3], act, 3, 2, 1))


def stem_pool_weight(K, pad_t, pad_l):
    """Host-side (tiny, load-time) re-ordering of an OIHW stride-2 first-layer filter into the [Cout][T][64] K-major
    layout of ``plnr_stem_pool_fwd``: W[co, e, (ph*3 + c)*8 + sx + col_shift] = K[co, c, 2(e + e_min) + ph + pad_t, sx]."""
    co, c, kh, kw = K.shape
    e_min, taps, shift = C.c_int(), C.c_int(), C.c_int()
    _capi.check(B.lib().plnr_stem_pool_geometry(kh, pad_t, pad_l, C.byref(e_min), C.byref(taps), C.byref(shift)),
                'plnr_stem_pool_geometry')
    out = np.zeros((co, taps.value, 64), np.float16)
    for e in range(taps.value):
        for ph in range(2):
            r = 2 * (e + e_min.value) + ph + pad_t
            if 0 <= r < kh:
                for ci in range(c):
                    k0 = (ph * 3 + ci) * 8 + shift.value
                    out[:, e, k0:k0 + kw] = K[:, ci, r, :]
    return out


def stem_pool_into(x_flat, w_packed, scale, shift, y, kh, kw, stride, pads, act=ACT_RELU):
    n, c, h, w = x_flat.shape
    t = y.tensor()
    p = lambda a: a.ptr if a is not None else None
    if x_flat.dtype == np.uint8:                  # uint8 image: converted by the kernel's producer warps (exact)
        fn, what = B.lib().plnr_stem_pool_fwd_u8, 'plnr_stem_pool_fwd_u8'
    elif x_flat.dtype == np.float16:
        fn, what = B.lib().plnr_stem_pool_fwd, 'plnr_stem_pool_fwd'
    else:
        raise _capi.PlanerB200Error('stem_pool: the image must be float16 or uint8, got %s' % x_flat.dtype)
    _capi.check(fn(B.ctx(), x_flat.ptr, n, c, h, w, w_packed.ptr, p(scale), p(shift), kh, kw,
                   stride, pads[0], pads[1], pads[2], pads[3], act, 3, 2, 1, C.byref(t)), what)
    return y


def stem3x3_supported(dtype, c, cout, kh, kw, strides, dilations, pads):
    """True when the small-first-layer kernel (csrc/stem_direct.cu) applies: fp16 net, 3x3 / s1 / p1, c <= 3, cout <= 32."""
    if strides[0] != strides[1] or dilations[0] != dilations[1]:
        return False
    return bool(B.lib().plnr_stem3x3_supported(_capi.dtype_code(dtype), c, cout, kh, kw, strides[0], pads[0], pads[1], pads[2],
                                               pads[3], dilations[0]))


def stem3x3_host_filter(K16, scale, shift):
    """The small first-layer filter as the HOST arrays plnr_stem3x3_fwd takes (they travel in the kernel parameters): one
    read-back at load time."""
    host = lambda a, dt: None if a is None else np.ascontiguousarray(a.get() if isinstance(a, B.DeviceArray) else a, dtype=dt)
    return host(K16, np.float16), host(scale, np.float32), host(shift, np.float32)


def stem3x3_into(x_flat, K16, scale, shift, y, act=ACT_NONE, alpha=0.0):
    """x_flat: NCHW image (float16 / uint8) on the device; K16 (OIHW float16), scale, shift: numpy arrays from
    ``stem3x3_host_filter`` (device arrays are read back on every call); y: nhwc output."""
    n, c, h, w = x_flat.shape
    t = y.tensor()
    K16, scale, shift = stem3x3_host_filter(K16, scale, shift)
    p = lambda a: a.ctypes.data if a is not None else None
    _capi.check(B.lib().plnr_stem3x3_fwd(B.ctx(), x_flat.ptr, _capi.src_dtype_code(x_flat.dtype), n, c, h, w, p(K16), p(scale),
                                         p(shift), int(act), float(alpha), C.byref(t)), 'plnr_stem3x3_fwd')
    return y


def nhwc_to_nchw_into(x, y_flat):
    t = x.tensor()
    _capi.check(B.lib().plnr_nhwc_to_nchw(B.ctx(), C.byref(t), _capi.dtype_code(x.dtype), y_flat.ptr,
                                          _capi.dtype_code(y_flat.dtype)), 'plnr_nhwc_to_nchw')
    return y_flat
