"""GraphPlan -> device buffers, kernel launches and one CUDA graph (the replacement of the hot loop of
planer/net.py:43-70).

Per forward the host issues: one layout/cast kernel per graph input (NCHW -> pixel-major, reading the
caller's device array in place) and ONE ``plnr_graph_launch`` that replays every fused step and the
output transposes.  Weights are packed ([Cout][kh][kw][Cin]) and bias/BatchNorm folded into fp32
(scale, shift) vectors once, on the device, when the executor is built.
"""
import os

import numpy as np

from . import _capi, ops, plan as P
from . import backend as B
from .backend import DeviceArray


def _round_up(v, m):
    return (v + m - 1) // m * m


class Executor:
    def __init__(self, net, gplan, dtype, use_graph=True, input_dtypes=None):
        self.net, self.plan, self.dtype = net, gplan, np.dtype(dtype)
        self.use_graph = use_graph
        # dtype each graph input arrives in (uint8 images and fp32 arrays are converted by the input-time kernel)
        self.input_dtypes = [np.dtype(d) for d in input_dtypes] if input_dtypes else [self.dtype] * len(gplan.inputs)
        self.values = gplan.values
        self.arr = {}             # root value id -> DeviceArray
        self.launches = []        # closures, in order
        self.kinds = []           # step op per closure (for timers / launch accounting)
        self.names = []           # fused layer names per closure
        self.graph = None
        self.input_ids = list(gplan.inputs)
        self.in_arrays, self.out_arrays, self.out_flat = [], [], []
        self._keep = []
        self.split_convs = 0     # float32 convolutions routed to the tensor pipe (fp16 split operands)
        self._in_stage = {}       # graph input -> fp16 staging array (fp32 images on the fused first layer)
        self.pack_hits = self.pack_misses = 0
        self.nchw_exits = 0       # graph outputs written as NCHW by the producing conv's epilogue
        self.pool_folds = {}      # root of a conv output -> (fp32 partial-sum array, h*w): GlobalAveragePool folded into that conv
        self._build()

    # ------------------------------------------------------------------------------------------
    def _root(self, vid):
        return P._root(self.values, vid)

    def _weight(self, vid):
        return self.net.weights[self.net.inits.index(self.values[vid].name)]

    def _storage_c(self, vid):
        v = self.values[vid]
        if v.kind == 'input' and len(v.shape) == 4:
            return self._input_cpad(vid)
        return self._out_cpad(vid)

    def _out_cpad(self, vid):
        """Stored channels of a conv output that only leaves the graph (YOLO heads: 255 channels): padded to a multiple of 8
        in fp16 so that rows stay 16-byte aligned and the conv epilogue keeps its vector path; the filter gets zero rows for
        the pad channels and the NHWC -> NCHW exit reads the logical channels only."""
        v = self.values[vid]
        c = v.shape[1]
        if self.dtype != np.float16 or len(v.shape) != 4 or c % 8 == 0 or not v.is_output:
            return c
        prod = [st for st in self.plan.steps if st.out == vid]
        users = [st for st in self.plan.steps if vid in [self._root(r) for r in st.reads()]]
        if len(prod) != 1 or prod[0].op != 'conv' or users or prod[0].shortcut is not None or prod[0].attrs.get('group', 1) != 1:
            return c
        if prod[0].res is not None:
            return c                 # a fused residual operand has the logical channel count: keep the unpadded path
        if self._fused_stem_of(self._root(prod[0].ins[0])) is not None or self._stem_of(self._root(prod[0].ins[0])) is not None:
            return c
        return _round_up(c, 8)

    def _nchw_exit_ok(self, st, x, K):
        v = self.values[self._root(st.out)]
        if self.dtype != np.float16 or os.environ.get('PLNR_NO_NCHW_EXIT') == '1' or not v.is_output or len(v.shape) != 4:
            return False
        if st.res is not None or st.shortcut is not None or st.attrs.get('group', 1) != 1 or st.attrs.get('flip'):
            return False
        root = self._root(st.out)
        if any(root in [self._root(r) for r in u.reads()] for u in self.plan.steps) or v.slice_of is not None:
            return False
        if self.stems.get(self._root(st.ins[0])) is not None or x.layout != 'nhwc':
            return False
        a = st.attrs
        if K.shape[2] == 1 and K.shape[3] == 1 and tuple(a['strides']) == (1, 1) and not any(a['pads']) and x.shape[1] % 64 == 0 \
                and os.environ.get('PLNR_PW_HEADS', '1') != '0':
            # pointwise heads (YOLOv3: 1x1 -> 255 channels): the resident-filter GEMM kernel (conv_pw.cu) into rows padded to
            # 256 channels + the register transposer of the exit (0.96 of the copy bandwidth) beat the NCHW-writing epilogue
            # of the shift kernel (102 + 48 + 43 us for the three heads at batch 32)
            return False
        return ops.conv2d_out_nchw_supported(x, v.shape, K.shape[2], K.shape[3], a['strides'], a['dilations'], a['pads'])

    def _is_side_operand(self, vid):
        """True when ``vid`` is also the residual / shortcut operand of some step (``out = x + conv(x)`` on a graph input): its
        stored layout must then stay the logical one -- no channel padding, no first-layer packing."""
        for st in self.plan.steps:
            if st.res is not None and self._root(st.res) == vid:
                return True
            if st.shortcut is not None and self._root(st.shortcut[0]) == vid:
                return True
        return False

    def _input_cpad(self, vid):
        """Graph inputs feeding only group-1 convs are channel-padded to a multiple of 16 in fp16 so that the
        first layer runs on the tensor cores (Cin=3 -> 16; the packed weights carry zeros there)."""
        c = self.values[vid].shape[1]
        if self.dtype != np.float16 or c % 16 == 0 or self.values[vid].is_output or self._is_side_operand(vid):
            return c                 # an input that is also returned (or added to a conv result) keeps its logical channels
        for st in self.plan.steps:
            if vid in [self._root(r) for r in st.reads()]:
                if not (st.op == 'conv' and st.attrs['group'] == 1 and self._root(st.ins[0]) == vid):
                    return c
        return _round_up(c, 16)

    def _stem_of(self, vid):
        """The conv step that can absorb graph input ``vid`` through ``plnr_stem_pack`` (few input channels, fp16):
        its filter taps along W, its stride phases along H and the input channels become the packed channels."""
        v = self.values[vid]
        if self.dtype != np.float16 or len(v.shape) != 4:
            return None
        users = [st for st in self.plan.steps if vid in [self._root(r) for r in st.reads()]]
        if len(users) != 1 or users[0].op != 'conv' or self._root(users[0].ins[0]) != vid or v.is_output:
            return None
        if self._is_side_operand(vid):
            return None
        st, a = users[0], users[0].attrs
        kshape = self.values[st.w].shape
        s = a['strides'][0]
        if a['group'] != 1 or a['dilations'] != (1, 1) or a['strides'][0] != a['strides'][1] or s not in (1, 2):
            return None
        if s * kshape[3] * v.shape[1] > 64 or v.shape[1] % 16 == 0:
            return None
        return st

    def _direct_stem_of(self, vid):
        """The conv step when graph input ``vid`` feeds ONLY a 3x3 / stride-1 / pad-1 convolution with <= 3 input and <= 32
        output channels (YOLOv3's first layer): csrc/stem_direct.cu computes conv + scale / shift + activation on the CUDA
        cores straight from the caller's NCHW (fp16 / uint8) array.  None otherwise."""
        v = self.values[vid]
        if self.dtype != np.float16 or len(v.shape) != 4 or v.is_output or os.environ.get('PLNR_NO_DIRECT_STEM') == '1':
            return None
        users = [st for st in self.plan.steps if vid in [self._root(r) for r in st.reads()]]
        if len(users) != 1 or users[0].op != 'conv' or self._root(users[0].ins[0]) != vid:
            return None
        st, a = users[0], users[0].attrs
        if st.res is not None or st.shortcut is not None or a.get('flip') or a['group'] != 1:
            return None
        out = self.values[self._root(st.out)]
        if out.is_output or out.slice_of is not None:
            return None
        k = self.values[st.w].shape
        if not ops.stem3x3_supported(self.dtype, v.shape[1], k[0], k[2], k[3], a['strides'], a['dilations'], a['pads']):
            return None
        return st

    def _fused_stem_of(self, vid):
        """(conv step, maxpool step) when graph input ``vid`` feeds conv -> (bn) -> relu -> maxpool(3x3/s2/p1) and the
        fused first-layer kernel (csrc/stem_pool.cu) supports the shapes; else None."""
        st = self._stem_of(vid)
        if st is None or st.res is not None or os.environ.get('PLNR_NO_FUSED_STEM') == '1':
            return None
        out = self._root(st.out)
        users = [u for u in self.plan.steps if out in [self._root(r) for r in u.reads()]]
        if len(users) != 1 or users[0].op != 'maxpool' or self.values[st.out].is_output:
            return None
        pool, v, a = users[0], self.values[vid], st.attrs
        kshape = self.values[st.w].shape
        if x_dtype_ok(self.dtype) and ops.stem_pool_supported(
                self.dtype, v.shape[1], v.shape[2], v.shape[3], kshape[0], kshape[2], kshape[3], a['strides'][0],
                a['pads'], st.act, pool.attrs['w'], pool.attrs['strides'], pool.attrs['pads']):
            return st, pool
        return None

    @staticmethod
    def _affine_host(bias, bn_k, bn_b, co):
        """(scale, shift) of a conv epilogue on the host, fp32: y = acc * scale + shift  (planer/layer.py:26, :125-127)."""
        k = np.ones(co, np.float32) if bn_k is None else bn_k.get().astype(np.float32).reshape(-1)
        b = np.zeros(co, np.float32) if bn_b is None else bn_b.get().astype(np.float32).reshape(-1)
        bi = np.zeros(co, np.float32) if bias is None else bias.get().astype(np.float32).reshape(-1)
        return k, bi * k + b

    def _step_affine(self, st, co):
        bias = self._weight(st.bias) if st.bias is not None else None
        bn_k, bn_b = (self._weight(st.bn[0]), self._weight(st.bn[1])) if st.bn else (None, None)
        return self._affine_host(bias, bn_k, bn_b, co)

    def _shortcut_eligible(self, main, short):
        """Can ``short`` (1x1 conv on the shortcut branch) be accumulated inside ``main``'s kernel?  Shapes must suit the
        shift kernel, and folding the two BatchNorm scales into the shortcut's fp16 weights must be benign."""
        vals, a = self.values, main.attrs
        xs, x2s, ys = vals[main.ins[0]].shape, vals[short.ins[0]].shape, vals[main.out].shape
        if len(xs) != 4 or len(x2s) != 4 or a['group'] != 1 or vals[short.ins[0]].kind == 'input':
            return False
        kshape = vals[main.w].shape
        if not ops.conv2d_shortcut_supported(self.dtype, xs, x2s, short.attrs['strides'][0], ys, kshape[2], kshape[3],
                                             a['strides'], a['dilations'], a['pads']):
            return False
        s_main, _ = self._step_affine(main, kshape[0])
        s_short, _ = self._step_affine(short, kshape[0])
        if np.any(np.abs(s_main) < 1e-3 * max(np.abs(s_main).max(), 1e-30)):
            return False
        return bool(np.abs(s_short / s_main).max() <= 64.0)

    def _gap_dense_tail(self, gap):
        """The dense step of a  gap -> flatten -> dense  tail (each value used once, none a graph output), or None."""
        if os.environ.get('PLNR_NO_GAP_DENSE') == '1' or len(self.values[gap.ins[0]].shape) != 4:
            return None
        cur, steps, chain = gap, self.plan.steps, []
        for want in ('flatten', 'dense'):
            if self.values[cur.out].is_output:
                return None
            users = [u for u in steps if cur.out in u.reads() and u is not cur]
            if len(users) != 1 or users[0].op != want:
                return None
            cur = users[0]
            chain.append(cur)
        if cur.res is not None or self._root(cur.ins[0]) != self._root(gap.out):
            return None
        c = self.values[gap.ins[0]].shape[1]
        if c % (16 // self.dtype.itemsize) != 0 or c * 8 * 4 > 96 * 1024:
            return None
        return chain

    def _pool_fold_of(self, st, x, K):
        """GlobalAveragePool folded into the epilogue of the convolution that feeds it (plnr_epilogue.pool_sum): applies when
        the conv's output is read by ONE step, a gap that heads a gap -> flatten -> dense tail, is not a graph output, and the
        library says the kernel it will run can pool (fp16 stride-1 shift GEMM, image grid of whole 32-position parts -- the
        7x7 map of a 224x224 ResNet).  Returns the parts per image or 0."""
        if self.dtype != np.float16 or st.attrs.get('group', 1) != 1 or st.shortcut is not None or self.values[st.out].is_output:
            return 0
        if os.environ.get('PLNR_NO_POOL_FOLD') == '1' or self.values[self._root(st.out)].slice_of is not None:
            return 0
        root = self._root(st.out)
        users = [u for u in self.plan.steps if u is not st and root in [self._root(r) for r in u.reads()]]
        if len(users) != 1 or users[0].op != 'gap' or self._gap_dense_tail(users[0]) is None:
            return 0
        a = st.attrs
        return ops.conv2d_pool_parts(x, self.values[st.out].shape, K.shape[2], K.shape[3], a['strides'], a['dilations'], a['pads'])

    def _packed(self, key, build):
        """One-off weight artefact of this net (packed filter, folded scale / shift, ...): taken from the Net's pack store when
        an executor of another input shape -- or the on-disk pack cache (io.py) -- already made it, else built and kept
        there.  ``build`` returns a DeviceArray or a tuple of DeviceArrays / None."""
        store = self.net._pack_store
        key = '%s|%s' % (self.dtype.name, key)
        if key + '#n' in store:
            n = store[key + '#n']
            vals = tuple(store.get('%s#%d' % (key, i)) for i in range(abs(n)))
            self.pack_hits += 1
            return vals if n > 0 else vals[0]
        val = build()
        vals = val if isinstance(val, tuple) else (val,)
        for i, v in enumerate(vals):
            if v is not None:
                store['%s#%d' % (key, i)] = v
        store[key + '#n'] = len(vals) if isinstance(val, tuple) else -1
        self.pack_misses += 1
        return val

    def _nbytes(self, vid, itemsize=None):
        shp = self.values[vid].shape
        return int(np.prod(shp)) * (self.dtype.itemsize if itemsize is None else itemsize)

    def _step_bytes(self, st, extra_weights=()):
        """Algorithmic bytes of one fused launch: logical operands read once, result written once."""
        n = sum(self._nbytes(r) for r in st.reads()) + self._nbytes(st.out)
        for w in (st.w, st.bias) + tuple(st.bn or ()) + tuple(extra_weights):
            if w is not None:
                n += self._nbytes(w)
        if st.shortcut is not None:
            n += self._nbytes(st.shortcut[2].w)
        return n

    def algorithmic_bytes(self):
        """Algorithmic HBM bytes of one forward as this executor fuses it: input-time kernels + every graph launch."""
        return int(self.input_bytes + sum(self.launch_bytes))

    def _get(self, vid):
        return self.arr[self._root(vid)]

    def _view(self, vid):
        """DeviceArray of value ``vid`` with ITS logical shape on its root's storage."""
        a, v = self._get(vid), self.values[vid]
        if a.shape == v.shape or self._root(vid) in getattr(self, 'stems', {}):
            return a
        if a.layout == 'nhwc' and len(v.shape) == 2:          # flatten of an (N,C,1,1) map
            return DeviceArray(a.buf, v.shape, a.dtype, 'flat', offset=a.offset)
        if a.layout == 'nhwc' and len(v.shape) == 4 and a.shape[0] == v.shape[0] and a.shape[2:] == v.shape[2:] \
                and a.shape[1] >= v.shape[1] and v.kind == 'input':
            return a                                          # channel-padded graph input (zeros beyond C)
        raise AssertionError('alias with a different shape: %s vs %s' % (a.shape, v.shape))

    def _build(self):
        gp, vals, dt = self.plan, self.values, self.dtype
        if dt == np.float16 and os.environ.get('PLNR_NO_SHORTCUT_FUSION') != '1':
            P.absorb_shortcuts(gp, self._shortcut_eligible)
        self.placed_concat_inputs = 0
        if os.environ.get('PLNR_NO_ZERO_COPY_CONCAT') != '1':
            self.placed_concat_inputs = P.place_concat_inputs(gp, align=16 // dt.itemsize)
        P.assign_buffers(gp, dt.itemsize, self._storage_c)
        pool = [None] * len(gp.buffer_bytes)

        def alloc(vid):
            r = self._root(vid)
            if r in self.arr:
                return self.arr[r]
            if vals[r].slice_of is not None:              # zero-copy concat: this value lives in a slice of the concat's buffer
                out, off = vals[r].slice_of
                self.arr[r] = ops.channel_slice(alloc(out), off, vals[r].shape[1])
                return self.arr[r]
            shape = vals[r].shape
            b = gp.buffer_of[r]
            if pool[b] is None:
                pool[b] = B.empty((gp.buffer_bytes[b],), np.uint8).buf
            layout = 'nhwc' if len(shape) == 4 else 'flat'
            ld = self._out_cpad(r) if layout == 'nhwc' else None
            self.arr[r] = DeviceArray(pool[b], shape, dt, layout, ld=ld)
            return self.arr[r]

        # graph inputs: pixel-major staging filled by an eager transform at every forward
        self.stems = {}
        self.fused_dense = set()
        self.fused_stems = {}     # graph input -> dict(conv, pool, run): conv+bn+relu+maxpool in ONE kernel at input time
        for vid in self.input_ids:
            shp = vals[vid].shape
            fs = self._fused_stem_of(vid)
            ds = self._direct_stem_of(vid) if fs is None else None
            st = self._stem_of(vid)
            if fs is not None:
                self.fused_stems[vid] = dict(conv=fs[0], pool=fs[1], run=None)
                a = None                                   # the kernel reads the caller's NCHW array in place
            elif ds is not None:
                self.fused_stems[vid] = dict(conv=ds, pool=None, run=None)       # small first layer on the CUDA cores, at input time
                a = None
            elif st is not None:
                kshape, at = vals[st.w].shape, st.attrs
                g = ops.stem_geometry(shp[2], shp[3], kshape[2], kshape[3], at['strides'][0], at['pads'])
                real = at['strides'][0] * kshape[3] * shp[1]
                cp = 16 if real <= 16 else (32 if real <= 32 else 64)
                a = B.empty((shp[0], cp, g['h2'], g['ow']), dt, 'nhwc')
                self.stems[vid] = dict(step=st, geom=g, cp=cp, kw=kshape[3], stride=at['strides'][0], pad_l=at['pads'][1])
            elif len(shp) == 4:
                cp = self._input_cpad(vid)
                a = B.empty((shp[0], cp, shp[2], shp[3]), dt, 'nhwc')
            else:
                a = B.empty(shp, dt)
            self.arr[vid] = a
            self.in_arrays.append(a)

        self.launch_bytes = []    # algorithmic bytes per launch: every operand read once + the result written once
        for st in gp.steps:
            self._io_override = None
            fn = self._make(st, alloc)
            if fn is not None:
                self.launches.append(fn)
                self.kinds.append(st.op)
                self.names.append(getattr(self, 'names_override', None) or '+'.join(st.fused))
                self.names_override = None
                self.launch_bytes.append(self._io_override if self._io_override is not None else self._step_bytes(st))
        self.input_bytes = 0      # input-time kernels (fused first layer / layout): graph input in its own dtype + what they write
        for k, vid in enumerate(self.input_ids):
            self.input_bytes += self._nbytes(vid, self.input_dtypes[k].itemsize)
            fs = self.fused_stems.get(vid)
            if fs is not None:
                self.input_bytes += self._nbytes((fs['pool'] or fs['conv']).out) + self._nbytes(fs['conv'].w)
            elif self.arr.get(vid) is not None:
                a = self.arr[vid]
                self.input_bytes += int(np.prod(a.shape)) * a.dtype.itemsize

        # graph outputs: restore NCHW (planer/net.py:100 hands NCHW arrays back)
        for o in gp.outputs:
            a = self._view(o)
            if a.layout == 'nhwc':
                flat = B.empty(a.shape, dt)
                self.launches.append(lambda a=a, flat=flat: ops.nhwc_to_nchw_into(a, flat))
                self.kinds.append('to_nchw')
                self.names.append('to_nchw')
                self.launch_bytes.append(2 * self._nbytes(o))
                self.out_flat.append(flat)
            else:
                self.out_flat.append(a)

    # ------------------------------------------------------------------------------------------
    def _make(self, st, alloc):
        vals, dt = self.values, self.dtype
        op = st.op
        if id(st) in self.fused_dense:
            return None                                  # computed by the gap step's fused kernel
        if op in ('conv', 'dense'):
            fused = self.fused_stems.get(self._root(st.ins[0])) if op == 'conv' else None
            fused = fused if fused is not None and fused['conv'] is st else None
            x = self._view(st.ins[0]) if fused is None else None
            K = self._weight(st.w)
            if op == 'conv' and st.attrs.get('flip'):
                K0 = K                                   # convtranspose: (C_in, C_out, kh, kw) -> flipped (C_out, C_in, kh, kw)
                K = self._packed(st.name + '|flip', lambda: ops.flip_weight(K0))
                self._keep.append(K)
            co = K.shape[0]
            bias = self._weight(st.bias) if st.bias is not None else None
            bn_k, bn_b = (self._weight(st.bn[0]), self._weight(st.bn[1])) if st.bn else (None, None)
            scale = shift = None
            if bias is not None or bn_k is not None:
                scale, shift = self._packed(st.name + '|fold', lambda: ops.fold_affine(bias, bn_k, bn_b, co))
                if bn_k is None:
                    scale = None
            res = self._view(st.res) if st.res is not None else None
            self._keep += [scale, shift]
            if fused is None and op == 'conv' and self._nchw_exit_ok(st, x, K):
                # graph output produced by this conv and read by nothing else: the epilogue writes the dense NCHW array
                # itself (plnr_epilogue.out_nchw) -- no pixel-major copy, no transpose launch, no channel padding
                a = st.attrs
                flat = B.empty(vals[st.out].shape, dt)
                self.arr[self._root(st.out)] = flat
                wp = self._packed('%s|pack|%d' % (st.name, x.shape[1]), lambda: ops.pack_weight(K, x.shape[1], dt))
                self._keep += [wp, flat]
                kh, kw = K.shape[2], K.shape[3]
                self.nchw_exits += 1
                return lambda: ops.conv2d_into(x, wp, flat, kh, kw, a['strides'], a['dilations'], a['pads'], 1, scale, shift,
                                               None, st.act, st.alpha, out_nchw=True)
            if fused is not None and fused['pool'] is None:
                y = alloc(st.out)
                K16 = self._packed(st.name + '|cast', lambda: K.astype(np.float16))
                self._keep += [K16, y]
                hK, hs, hf = ops.stem3x3_host_filter(K16, scale, shift)      # 1.7 KB, travel in the kernel parameters
                fused['run'] = lambda xf: ops.stem3x3_into(xf, hK, hs, hf, y, st.act, st.alpha)
                return None
            if fused is not None:
                a = st.attrs
                yp = alloc(fused['pool'].out)
                wp = self._packed(st.name + '|stem_pool|%d|%d' % (a['pads'][0], a['pads'][1]),
                                  lambda: B.asarray(ops.stem_pool_weight(K.get().astype(np.float16), a['pads'][0], a['pads'][1])))
                self._keep.append(wp)
                kh, kw = K.shape[2], K.shape[3]
                fused['run'] = lambda xf: ops.stem_pool_into(xf, wp, scale, shift, yp, kh, kw, a['strides'][0], a['pads'],
                                                            st.act)
                return None
            y = alloc(st.out)
            if op == 'conv' and st.shortcut is not None:
                # conv + bn + add(bn_d(conv1x1_d(x2))) + act in ONE launch: the shortcut's weights, scaled by the ratio of
                # the two BatchNorm scales, are appended to the packed filter along K; the shifts add up
                x2_vid, s2, sh = st.shortcut
                x2 = self._view(x2_vid)
                a = st.attrs
                def build_shortcut():
                    Kd = self._weight(sh.w).get().astype(np.float32)[:, :, 0, 0]
                    s_main, t_main = self._step_affine(st, co)
                    s_short, t_short = self._step_affine(sh, co)
                    w2 = (Kd * (s_short / s_main)[:, None]).astype(np.float16)
                    wp = ops.pack_weight(K, x.shape[1], dt).get().reshape(co, -1)
                    return (B.asarray(np.ascontiguousarray(np.concatenate([wp, w2], axis=1))),
                            B.asarray(s_main.astype(np.float32)), B.asarray((t_main + t_short).astype(np.float32)))
                wcat, scale_c, shift_c = self._packed('%s|shortcut|%s|%d' % (st.name, sh.name, x.shape[1]), build_shortcut)
                self._keep += [wcat, scale_c, shift_c]
                kh, kw = K.shape[2], K.shape[3]
                return lambda: ops.conv2d_shortcut_into(x, wcat, x2, s2, y, kh, kw, a['strides'], a['dilations'], a['pads'],
                                                        scale_c, shift_c, st.act, st.alpha)
            stem = self.stems.get(self._root(st.ins[0])) if op == 'conv' else None
            if stem is not None:
                # first layer on the packed input: (T x 1) stride-1 conv, taps re-ordered on the host (tiny, load time)
                g = stem['geom']
                wp = self._packed('%s|stem_pack|%d|%s|%d' % (st.name, stem['stride'], tuple(st.attrs['pads']), stem['cp']),
                                  lambda: B.asarray(ops.stem_pack_weight(K.get().astype(np.float16), stem['stride'],
                                                                         st.attrs['pads'], stem['cp'])))
                self._keep.append(wp)
                pads2 = (g['pad_t2'], 0, g['pad_b2'], 0)
                return lambda: ops.conv2d_into(x, wp, y, g['T'], 1, (1, 1), (1, 1), pads2, 1, scale, shift, res,
                                               st.act, st.alpha, res_after_act=st.res_after)
            if op == 'conv':
                a = st.attrs
                g = a['group']
                kh, kw = K.shape[2], K.shape[3]
                if y.layout == 'nhwc' and y.ld != y.shape[1] and vals[self._root(st.out)].slice_of is None:
                    # channel-padded graph output (_out_cpad): run the kernel on ld channels, pad filter rows are zero
                    cop = y.ld
                    wp = self._packed('%s|pack|%d|%d' % (st.name, x.shape[1], cop), lambda: ops.pack_weight(K, x.shape[1], dt, co_pad=cop))
                    pad1 = lambda v: None if v is None else ops.pad_vector(v, cop)
                    scale, shift = self._packed('%s|foldpad|%d' % (st.name, cop), lambda: (pad1(scale), pad1(shift)))
                    y = DeviceArray(y.buf, (y.shape[0], cop) + y.shape[2:], dt, 'nhwc', ld=cop, offset=y.offset)
                    self._keep += [wp, scale, shift]
                    return lambda: ops.conv2d_into(x, wp, y, kh, kw, a['strides'], a['dilations'], a['pads'], 1,
                                                   scale, shift, res, st.act, st.alpha, res_after_act=st.res_after)
                if dt == np.float32 and K.dtype == np.float32 and g == 1 and ops.split_conv_enabled():
                    # float32 on the tensor pipe: fp16 (hi, lo) split operands, fp32 accumulator and epilogue (csrc/split_f32.cu)
                    w16, meta = self._packed('%s|split|%d' % (st.name, x.shape[1]), lambda: ops.pack_weight_split(K))
                    sw = ops.split_weight(w16, meta, x.shape[1])
                    self._keep.append(sw)
                    self.split_convs += 1
                    return lambda: ops.conv2d_into(x, sw, y, kh, kw, a['strides'], a['dilations'], a['pads'], 1,
                                                   scale, shift, res, st.act, st.alpha, res_after_act=st.res_after)
                wp = self._packed('%s|pack|%d' % (st.name, x.shape[1] // g), lambda: ops.pack_weight(K, x.shape[1] // g, dt))
                self._keep.append(wp)
                parts = self._pool_fold_of(st, x, K)
                if parts > 0:
                    # the only reader is gap -> flatten -> dense: the epilogue writes per-part sums instead of the activation
                    n_, c_, h_, w_ = vals[st.out].shape
                    pool = B.empty((n_, parts, c_), np.float32)
                    self.pool_folds[self._root(st.out)] = (pool, h_ * w_)
                    self._keep.append(pool)
                    return lambda: ops.conv2d_into(x, wp, y, kh, kw, a['strides'], a['dilations'], a['pads'], g,
                                                   scale, shift, res, st.act, st.alpha, res_after_act=st.res_after, pool_sum=pool)
                return lambda: ops.conv2d_into(x, wp, y, kh, kw, a['strides'], a['dilations'], a['pads'], g,
                                               scale, shift, res, st.act, st.alpha, res_after_act=st.res_after)
            Kc = self._packed(st.name + '|cast', lambda: _aligned_cast(K, dt))
            self._keep.append(Kc)
            if x.layout != 'flat':
                raise NotImplementedError('dense %r needs a 2-D input (got %s)' % (st.name, x.shape))
            return lambda: ops.dense_into(x, Kc, y, scale, shift, res, st.act, st.alpha, res_after_act=st.res_after)
        if op == 'relu':
            x = self._get(st.ins[0])                      # in place on the root storage
            if self._root(st.ins[0]) not in self.arr:
                raise AssertionError('relu input not materialised')
            return lambda: ops.eltwise(ops.EW_RELU, _dense(x), _dense(x))
        if op == 'alias':
            return None
        if op in ('leakyrelu', 'sigmoid'):
            x, y = self._view(st.ins[0]), alloc(st.out)
            code = ops.EW_LEAKY if op == 'leakyrelu' else ops.EW_SIGMOID
            alpha = st.attrs.get('alpha', 0.0)
            return lambda: ops.eltwise(code, _dense(x), y, alpha=alpha)
        if op == 'add':
            x1, x2, y = self._view(st.ins[0]), self._view(st.ins[1]), alloc(st.out)
            return lambda: ops.eltwise(ops.EW_ADD, _dense(x1), y, p0=_dense(x2))
        if op == 'scale_shift':
            x, y = self._view(st.ins[0]), alloc(st.out)
            k, b = self._weight(st.bn[0]).astype(dt), self._weight(st.bn[1]).astype(dt)
            self._keep += [k, b]
            return lambda: ops.eltwise(ops.EW_SCALE_SHIFT, _dense(x), y, p0=k, p1=b)
        if op == 'maxpool':
            if any(f['pool'] is st for f in self.fused_stems.values()):
                return None                              # computed by the fused first-layer kernel
            x, y, a = self._view(st.ins[0]), alloc(st.out), st.attrs
            return lambda: ops.maxpool_into(x, y, a['w'], a['pads'], a['strides'])
        if op == 'upsample':
            x, y, a = self._view(st.ins[0]), alloc(st.out), st.attrs
            # the interpolation tables are uploaded now (outside graph capture) and OWNED by this executor: its CUDA graph
            # keeps their addresses
            if a.get('mode') == 'linear_size':
                tabs = ops.resize_linear_device_tables(x, y)
                self._keep.append(tabs)
                return lambda: ops.resize_linear_into(x, y, tabs)
            if a.get('mode', 'nearest') == 'linear':
                wm = ops.upsample_linear_weights(a['fh'], a['fw'])
                self._keep.append(wm)
                return lambda: ops.upsample_linear_into(x, y, a['fh'], a['fw'], wm)
            return lambda: ops.upsample_into(x, y, a['fh'], a['fw'])
        if op == 'clip':
            x, a = self._get(st.ins[0]), st.attrs         # in place on the root storage, like relu (planer/layer.py:250-251)
            if self._root(st.ins[0]) not in self.arr:
                raise AssertionError('clip input not materialised')
            return lambda: ops.unary2(ops.EW_CLIP, _dense(x), _dense(x), a.get('min', 0), a.get('max', 1))
        if op == 'hardsigmoid':
            x, y, a = self._view(st.ins[0]), alloc(st.out), st.attrs
            return lambda: ops.unary2(ops.EW_HARDSIGMOID, _dense(x), y, a.get('alpha', 0.2), a.get('beta', 0.5))
        if op == 'softmax':
            x, y = self._view(st.ins[0]), alloc(st.out)
            return lambda: ops.softmax_into(_dense(x), y)
        if op == 'averagepool':
            x, y, a = self._view(st.ins[0]), alloc(st.out), st.attrs
            return lambda: ops.avgpool_into(x, y, a['w'], a['pads'], a['strides'])
        if op == 'zero_stuff':
            x, y, a = self._view(st.ins[0]), alloc(st.out), st.attrs
            return lambda: ops.zero_stuff_into(x, y, a['lo_h'], a['lo_w'], a['strides'])
        if op == 'concat':
            y = alloc(st.out)
            xs = [self._view(i) for i in st.ins]
            offs = np.cumsum([0] + [x.shape[1] for x in xs]).tolist()
            out_root = self._root(st.out)
            # inputs the planner placed inside y (P.place_concat_inputs) are already there; copy only the others
            todo = [(x, ops.channel_slice(y, o, x.shape[1])) for x, o, i in zip(xs, offs, st.ins)
                    if vals[self._root(i)].slice_of != (out_root, o)]
            if not todo:
                return None
            def run():
                for x, v in todo:
                    ops.copy_channels(x, v)
            return run
        if op == 'gap':
            tail = self._gap_dense_tail(st)
            if tail is not None:
                # gap -> flatten -> dense: ONE small kernel instead of a pooling launch + a tensor-core GEMM launch
                fl, dn = tail
                x = self._view(st.ins[0])
                K = self._weight(dn.w)
                bias = self._weight(dn.bias) if dn.bias is not None else None
                bn_k, bn_b = (self._weight(dn.bn[0]), self._weight(dn.bn[1])) if dn.bn else (None, None)
                scale = shift = None
                if bias is not None or bn_k is not None:
                    scale, shift = self._packed(dn.name + '|fold', lambda: ops.fold_affine(bias, bn_k, bn_b, K.shape[0]))
                    if bn_k is None:
                        scale = None
                Kc = self._packed(dn.name + '|cast', lambda: _aligned_cast(K, dt))
                y = alloc(dn.out)
                self._keep += [scale, shift, Kc]
                self.fused_dense |= {id(fl), id(dn)}
                self.names_override = '+'.join(st.fused + dn.fused)
                self._io_override = self._nbytes(st.ins[0]) + self._nbytes(dn.out) + self._nbytes(dn.w)
                fold = self.pool_folds.get(self._root(st.ins[0]))
                if fold is not None:
                    pool, hw = fold
                    self._io_override = pool.nbytes + self._nbytes(dn.out) + self._nbytes(dn.w)
                    return lambda: ops.pooled_dense_into(pool, hw, Kc, y, scale, shift, dn.act, dn.alpha)
                return lambda: ops.gap_dense_into(x, Kc, y, scale, shift, dn.act, dn.alpha)
            x, y = self._view(st.ins[0]), alloc(st.out)
            return lambda: ops.gap_into(x, y)
        if op == 'flatten':
            x = self._view(st.ins[0])
            if self.values[st.out].alias_of is not None:
                return None                              # (N,C,1,1) -> (N,C): same memory
            y = alloc(st.out)
            flat4 = DeviceArray(y.buf, x.shape, dt, 'flat', offset=y.offset)
            return lambda: ops.nhwc_to_nchw_into(x, flat4)
        raise NotImplementedError(op)

    # ------------------------------------------------------------------------------------------
    def _load_inputs(self, xs):
        for k, (a, x, vid) in enumerate(zip(self.in_arrays, xs, self.input_ids)):
            shp = self.values[vid].shape
            if tuple(x.shape) != tuple(shp):
                raise ValueError('input %r: plan compiled for shape %s, got %s' % (self.values[vid].name, shp, x.shape))
            if x.dtype != self.input_dtypes[k]:
                raise ValueError('input %r: plan compiled for dtype %s, got %s' % (self.values[vid].name,
                                                                                  self.input_dtypes[k], x.dtype))
            if x.layout != 'flat':
                x = B.to_flat(x)
            if vid in self.fused_stems:
                if x.dtype not in (np.float16, np.uint8):           # fp32 host image on an fp16 net: one cast, then the fused kernel
                    stage = self._in_stage.setdefault(vid, B.empty(shp, np.float16))
                    _capi.check(B.lib().plnr_cast(B.ctx(), x.ptr, _capi.src_dtype_code(x.dtype), stage.ptr, _capi.F16, x.size),
                                'plnr_cast')
                    x = stage
                self.fused_stems[vid]['run'](x)
            elif vid in self.stems:
                sm = self.stems[vid]
                ops.stem_pack_into(x, a, sm['kw'], sm['stride'], sm['pad_l'])
            elif len(shp) == 4:
                ops.nchw_to_nhwc_into(x, a, shp[1])
            else:
                _capi.check(B.lib().plnr_cast(B.ctx(), x.ptr, _capi.src_dtype_code(x.dtype), a.ptr,
                                              _capi.dtype_code(a.dtype), x.size), 'plnr_cast')

    def run(self, xs):
        """xs: DeviceArrays (flat NCHW) in graph-input order -> tuple of flat (NCHW) output DeviceArrays that
        stay owned by the executor (valid until the next run)."""
        self._load_inputs(xs)
        if not self.use_graph:
            for fn in self.launches:
                fn()
        else:
            import ctypes as C
            if self.graph is None:
                for fn, name in zip(self.launches, self.names):        # one eager pass first: validates every launch outside capture
                    try:
                        fn()
                    except _capi.PlanerB200Error as e:
                        raise _capi.PlanerB200Error('step %r: %s' % (name, e)) from None
                B.synchronize()
                lib, ctx = B.lib(), B.ctx()
                _capi.check(lib.plnr_graph_begin(ctx), 'plnr_graph_begin')
                try:
                    for fn in self.launches:
                        fn()
                finally:
                    g = C.c_void_p()
                    rc = lib.plnr_graph_end(ctx, C.byref(g))
                _capi.check(rc, 'plnr_graph_end')
                self.graph = g
                # this forward was already computed by the eager pass; replaying the graph now would read first-layer
                # outputs (written at input time, outside the graph) that later layers have since recycled
                return tuple(self.out_flat)
            _capi.check(B.lib().plnr_graph_launch(B.ctx(), self.graph), 'plnr_graph_launch')
        return tuple(self.out_flat)

    def close(self):
        if self.graph is not None:
            B.lib().plnr_graph_destroy(self.graph)
            self.graph = None


def x_dtype_ok(dt):
    return np.dtype(dt) == np.float16


def _aligned_cast(K, dt):
    """K in the compute dtype at a 16-byte-aligned address: ``astype`` of an array that already has the dtype returns the
    array itself, i.e. a view into the weight blob at whatever byte offset the model file gave it (planer/net.py:83-88 packs
    the inits back to back) -- the vector loads of the dense kernels need alignment."""
    Kc = K.astype(dt)
    return B.clone(Kc) if Kc.ptr % 16 else Kc


def _dense(a):
    if a.layout == 'nhwc':
        assert a.ld == a.shape[1] and a.coff == 0
    return a
