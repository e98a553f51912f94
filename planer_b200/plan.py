"""Flow compiler: Planer's JSON IR -> a short list of fused kernel launches with static buffers.

The reference executes ``flow`` with a Python interpreter, one array-library call chain per layer
(planer/net.py:37-72): ResNet-18 = 70 dispatches per forward, about as long as the whole forward takes on a
B200.  Here the same flow is compiled once per input shape:

  1. ``flatten``      -- flow entries (incl. chained ones, net.py:46-50) -> single-op SSA nodes; the in-place
                         ReLU aliasing of the reference (layer.py:44-46) is kept by renaming
                         every key that points at the mutated value.
  2. ``infer``        -- shapes per value with the reference's formulas (util.py:25-26, :84-85).
  3. ``fuse``         -- conv/dense -> batchnorm -> add -> relu|leakyrelu|sigmoid chains collapse into the conv
                         epilogue when each intermediate has exactly one consumer and is not a graph output.
  4. ``assign``       -- liveness-based reuse of activation buffers.

Everything in this file is host logic without device access (CPU-testable); ``executor.py`` turns a GraphPlan
into kernel launches and a CUDA graph.  Operators outside the hot path raise NotImplementedError.
"""
import numpy as np

HOT_OPS = ('conv', 'dense', 'relu', 'leakyrelu', 'sigmoid', 'add', 'batchnorm', 'flatten', 'gap', 'concat',
           'maxpool', 'upsample', 'identity', 'return',
           'averagepool', 'convtranspose', 'hardsigmoid', 'clip', 'softmax', 'resize')          # the same kernels either side of the path
ACTS = {'relu': 1, 'leakyrelu': 2, 'sigmoid': 3}


def _out(n_in, pad_lo, pad_hi, k, dil, stride):
    return (n_in + pad_lo + pad_hi - (k - 1) * dil - 1 + stride) // stride


class Value:
    __slots__ = ('id', 'shape', 'kind', 'name', 'producer', 'uses', 'is_output', 'alias_of', 'slice_of')

    def __init__(self, vid, shape, kind, name):
        self.id, self.shape, self.kind, self.name = vid, None if shape is None else tuple(shape), kind, name
        self.producer, self.uses, self.is_output, self.alias_of = None, 0, False, None
        self.slice_of = None      # (concat output value id, first channel): this value LIVES in a channel slice of that buffer


class Node:
    __slots__ = ('name', 'kind', 'attrs', 'ins', 'outs', 'flops')

    def __init__(self, name, kind, attrs, ins, outs):
        self.name, self.kind, self.attrs, self.ins, self.outs, self.flops = name, kind, attrs, ins, outs, 0


def flatten(model, input_shapes):
    """IR -> (values, nodes, output value ids).  ``input_shapes``: {input name: NCHW shape}."""
    layers = {name: (kind, attrs) for name, kind, attrs in model['layers']}
    values, cur = [], {}

    def new(shape, kind, name):
        v = Value(len(values), shape, kind, name)
        values.append(v)
        return v

    for name, shape, dt in model['inits']:
        cur[name] = new(shape, 'weight', name).id
    for name in model['input']:
        if name in cur and name not in input_shapes:
            continue                      # old ONNX files list initializers as inputs
        if name not in input_shapes:
            raise KeyError('no shape given for graph input %r' % name)
        cur[name] = new(input_shapes[name], 'input', name).id
    nodes, outputs = [], None
    for xs, names, y in model['flow']:
        names = names if isinstance(names, list) else [names]
        for j, lname in enumerate(names):
            src = xs if j == 0 else y
            if isinstance(src, str):
                ins = [cur[src]]                                   # net.py:50: KeyError if missing
            else:
                ins = [None if k in ('', 'None') else cur.get(k) for k in src]   # net.py:39,49
            kind, attrs = layers[lname]
            if kind not in HOT_OPS:
                raise NotImplementedError("operator %r (layer %r) is not on the B200 hot path (SURVEY section 8); "
                                          "planer_b200 has no CPU fallback" % (kind, lname))
            if kind == 'return':
                outputs = [i for i in ins]
                cur[y if isinstance(y, str) else y[0]] = None
                nodes.append(Node(lname, kind, attrs, ins, []))
                continue
            if not isinstance(y, str):
                raise NotImplementedError('multi-output layer %r is not on the B200 hot path' % lname)
            if kind == 'convtranspose':
                # planer/layer.py:28-34: zero-stuffing, then an ordinary stride-1 conv with the flipped, transposed filter
                if (attrs.get('group', 1) or 1) != 1:
                    raise NotImplementedError('convtranspose %r: group > 1 is not supported' % lname)
                mid = new(None, 'act', y + '.stuffed')
                mid.producer = len(nodes)
                nodes.append(Node(lname + '.stuff', 'zero_stuff', dict(attrs), [ins[0], ins[1]], [mid.id]))
                conv_attrs = {'group': 1, 'strides': (1, 1), 'dilations': tuple(attrs.get('dilations') or (1, 1)),
                              'pads': (0, 0, 0, 0), 'flip': True}
                out = new(None, 'act', y)
                out.producer = len(nodes)
                nodes.append(Node(lname, 'conv', conv_attrs, [mid.id] + ins[1:], [out.id]))
                cur[y] = out.id
                continue
            out = new(None, 'act', y)
            out.producer = len(nodes)
            nodes.append(Node(lname, kind, dict(attrs), ins, [out.id]))
            if kind in ('relu', 'identity', 'clip'):
                # in-place (planer/layer.py:44-46 ReLU, :250-251 Clip): every key that named the input now sees the
                # mutated / same array
                out.alias_of = ins[0]
                for k, vid in list(cur.items()):
                    if vid == ins[0]:
                        cur[k] = out.id
            cur[y] = out.id
    if outputs is None:                                           # graphs without a 'return' layer: net.py:72
        last = model['flow'][-1][2]
        outputs = [cur[last if isinstance(last, str) else last[0]]]
    return values, nodes, outputs


def infer(values, nodes, host_consts=None):
    """Fill value shapes and per-node algorithmic FLOPs (conv/dense only)."""
    host_consts = host_consts or {}
    for nd in nodes:
        sh = [None if i is None else values[i].shape for i in nd.ins]
        k, a = nd.kind, nd.attrs
        if k == 'conv':
            (n, c, h, w), (co, cg, kh, kw) = sh[0], sh[1]
            if a.get('flip'):
                co, cg = cg, co                      # convtranspose filter is stored (C_in, C_out, kh, kw)
            g = a.get('group', 1) or 1
            st = a.get('strides') or (1, 1)
            dl = a.get('dilations') or (1, 1)
            pd = a.get('pads') or (0, 0, 0, 0)
            if c != cg * g:
                raise ValueError('conv %r: input has %d channels, weight expects %d x %d groups' % (nd.name, c, cg, g))
            oh, ow = _out(h, pd[0], pd[2], kh, dl[0], st[0]), _out(w, pd[1], pd[3], kw, dl[1], st[1])
            if pd[2] > pd[0] or pd[3] > pd[1]:
                raise ValueError('conv %r: pads %s with bottom>top or right>left are undefined in the reference '
                                 '(planer/util.py:4-10, SURVEY App. D Q1)' % (nd.name, list(pd)))
            out = (n, co, oh, ow)
            nd.flops = 2 * n * co * cg * kh * kw * oh * ow
            nd.attrs = dict(group=g, strides=tuple(st), dilations=tuple(dl), pads=tuple(pd), flip=bool(a.get('flip')))
        elif k == 'dense':
            (m, kk), (nn, k2) = sh[0], sh[1]
            if kk != k2:
                raise ValueError('dense %r: x is %s, K is %s' % (nd.name, sh[0], sh[1]))
            out = (m, nn)
            nd.flops = 2 * m * nn * kk
        elif k in ('relu', 'leakyrelu', 'sigmoid', 'batchnorm', 'identity', 'hardsigmoid', 'clip'):
            out = sh[0]
            if k == 'clip' and len(nd.ins) > 1:
                # bounds as inputs: the reference's interpreter hands every input to Clip(x, min, max) (planer/net.py:54-58), so an
                # IR with constant-init bounds is valid there; anything but constants is rejected instead of silently ignored
                nd.attrs = dict(nd.attrs)
                for key, vid in zip(('min', 'max'), nd.ins[1:3]):
                    name = values[vid].name
                    if name not in host_consts:
                        raise NotImplementedError('clip %r: bound %r must be a constant init' % (nd.name, name))
                    nd.attrs[key] = float(np.asarray(host_consts[name], np.float64).reshape(-1)[0])
        elif k == 'softmax':
            ax = a.get('axis', -1)
            if not ((len(sh[0]) == 4 and ax in (1, -3)) or (len(sh[0]) == 2 and ax in (1, -1))):
                raise NotImplementedError('softmax %r: only the channel axis of 4-D tensors / the last axis of 2-D tensors '
                                          '(got %d-D, axis %d)' % (nd.name, len(sh[0]), ax))
            out = sh[0]
        elif k == 'add':
            if sh[0] != sh[1]:
                raise NotImplementedError('add %r: broadcasting (%s + %s) is outside the B200 hot path'
                                          % (nd.name, sh[0], sh[1]))
            out = sh[0]
        elif k == 'zero_stuff':
            (n, c, h, w), (ci, co, kh, kw) = sh[0], sh[1]
            if c != ci:
                raise ValueError('convtranspose %r: input has %d channels, weight expects %d' % (nd.name, c, ci))
            st, dl = a.get('strides') or (2, 2), a.get('dilations') or (1, 1)
            pd, op_ = a.get('pads') or (0, 0, 0, 0), a.get('output_padding') or (0, 0)
            lo_h, hi_h = (kh - 1) * dl[0] - pd[0], (kh - 1) * dl[0] - pd[2] + op_[0]
            lo_w, hi_w = (kw - 1) * dl[1] - pd[1], (kw - 1) * dl[1] - pd[3] + op_[1]
            if min(lo_h, hi_h, lo_w, hi_w) < 0:
                raise NotImplementedError('convtranspose %r: pads larger than (k-1)*dilation are not supported' % nd.name)
            out = (n, c, (h - 1) * st[0] + lo_h + hi_h + 1, (w - 1) * st[1] + lo_w + hi_w + 1)
            nd.attrs = dict(lo_h=lo_h, lo_w=lo_w, strides=tuple(st))
        elif k in ('maxpool', 'averagepool'):
            n, c, h, w = sh[0]
            kw_, pd, st = a.get('w', (2, 2)), a.get('pads', (0, 0, 0, 0)), a.get('strides', (2, 2))
            out = (n, c, (h + pd[0] + pd[2] - kw_[0] + st[0]) // st[0], (w + pd[1] + pd[3] - kw_[1] + st[1]) // st[1])
            nd.attrs = dict(w=tuple(kw_), pads=tuple(pd), strides=tuple(st))
        elif k in ('upsample', 'resize'):
            mode = a.get('mode', 'nearest')
            sname = values[nd.ins[1 if k == 'upsample' else 2]].name if len(nd.ins) > (1 if k == 'upsample' else 2) and \
                nd.ins[1 if k == 'upsample' else 2] is not None else None
            scales = host_consts.get(sname)
            if scales is None:
                raise ValueError('%s %r: scales tensor %r must be a constant init' % (k, nd.name, sname))
            sc = np.asarray(scales, np.float64).reshape(-1)[-2:]
            if k == 'resize' and sc.size == 2 and np.any(sc != np.floor(sc)):
                if mode != 'linear':
                    raise NotImplementedError('resize %r: nearest with fractional scales is not implemented' % nd.name)
                n, c, h, w = sh[0]
                out = (n, c, int(round(float(sc[0]) * h)), int(round(float(sc[1]) * w)))      # planer/util.py:214
                nd.attrs = dict(mode='linear_size')
                values[nd.outs[0]].shape = tuple(int(v) for v in out)
                continue
            if k == 'resize' and sc.size != 2:
                raise NotImplementedError('resize %r: needs a 4-entry scales tensor' % nd.name)
            f = sc.astype(int).tolist()                                           # planer/layer.py:82
            if mode not in ('nearest', 'linear') or (mode == 'linear' and min(f) < 2):
                raise NotImplementedError("%s %r: mode %r with factors %s is not implemented" % (k, nd.name, mode, f))
            if k == 'resize' and mode == 'nearest':
                from .layer import nearest_shift
                tm, rm = a.get('coordinate_transformation_mode', 'half_pixel'), a.get('nearest_mode', 'round_prefer_floor')
                if nearest_shift(f[0], tm, rm) or nearest_shift(f[1], tm, rm):
                    raise NotImplementedError('resize %r: nearest with a non-zero pixel shift (%s, %s)' % (nd.name, tm, rm))
            n, c, h, w = sh[0]
            out = (n, c, h * int(f[0]), w * int(f[1]))
            nd.attrs = dict(fh=int(f[0]), fw=int(f[1]), mode=mode)
        elif k == 'concat':
            if a.get('axis', 0) != 1 or any(len(s) != 4 for s in sh):
                raise NotImplementedError('concat %r: only channel concat (axis=1) of 4-D tensors is on the hot path' % nd.name)
            n, _, h, w = sh[0]
            out = (n, sum(s[1] for s in sh), h, w)
        elif k == 'gap':
            out = (sh[0][0], sh[0][1], 1, 1)
        elif k == 'flatten':
            out = (sh[0][0], int(np.prod(sh[0][1:])))
        elif k == 'return':
            continue
        else:                                                       # pragma: no cover (guarded by flatten)
            raise NotImplementedError(k)
        values[nd.outs[0]].shape = tuple(int(v) for v in out)
    return values


def infer_shapes(model, input_shapes, host_consts=None):
    """Convenience wrapper: per-node output shapes and FLOPs of one forward."""
    if host_consts is None:
        host_consts = {n: np.array([1, 1, 2, 2], np.float32) for n, s, d in model['inits'] if n.endswith('.scales') or n == 'S'}
    values, nodes, outputs = flatten(model, input_shapes)
    infer(values, nodes, host_consts)
    return {'nodes': [{'name': n.name, 'kind': n.kind, 'flops': n.flops,
                       'out_shape': values[n.outs[0]].shape if n.outs else None} for n in nodes],
            'outputs': [values[o].shape for o in outputs]}


class Step:
    """One kernel launch (or a zero-cost alias) of the compiled forward."""
    __slots__ = ('op', 'name', 'ins', 'out', 'attrs', 'w', 'bias', 'bn', 'res', 'act', 'alpha', 'fused', 'inplace',
                 'res_after', 'shortcut')

    def __init__(self, op, name, ins, out, attrs=None):
        self.op, self.name, self.ins, self.out, self.attrs = op, name, list(ins), out, dict(attrs or {})
        self.w = self.bias = self.bn = self.res = None
        self.act, self.alpha, self.fused, self.inplace, self.res_after = 0, 0.0, [name], False, False
        self.shortcut = None      # (input value id, stride, absorbed 1x1 conv Step): see absorb_shortcuts

    def reads(self):
        extra = [self.shortcut[0]] if self.shortcut else []
        return [i for i in self.ins + [self.res] + extra if i is not None]


class GraphPlan:
    def __init__(self):
        self.values, self.nodes, self.steps, self.inputs, self.outputs = [], [], [], [], []
        self.buffer_of, self.buffer_bytes, self.flops = {}, [], 0

    def summary(self):
        from collections import Counter
        return dict(Counter(s.op for s in self.steps))


def fuse(values, nodes, outputs):
    """Nodes -> Steps with conv/dense epilogue fusion."""
    for v in values:
        v.uses = 0
    for nd in nodes:
        for i in nd.ins:
            if i is not None:
                values[i].uses += 1
    for o in outputs:
        values[o].is_output = True
    steps, open_steps = [], {}      # open_steps: value id -> Step whose epilogue can still absorb its consumer

    def absorbable(vid):
        v = values[vid]
        return vid in open_steps and v.uses == 1 and not v.is_output

    def emit(st):
        steps.append(st)
        return st

    for nd in nodes:
        k = nd.kind
        if k == 'return':
            continue
        out = nd.outs[0]
        if k in ('conv', 'dense'):
            st = emit(Step(k, nd.name, [nd.ins[0]], out, nd.attrs))
            st.w = nd.ins[1]
            st.bias = nd.ins[2] if len(nd.ins) > 2 else None
            open_steps[out] = st
            continue
        x = nd.ins[0]
        if k == 'batchnorm' and absorbable(x):
            st = open_steps[x]
            if st.bn is None and st.res is None and st.act == 0:
                st.bn = (nd.ins[1], nd.ins[2])
                st.out = out
                st.fused.append(nd.name)
                del open_steps[x]
                open_steps[out] = st
                continue
        if k == 'add':
            done = False
            for a, b in ((nd.ins[0], nd.ins[1]), (nd.ins[1], nd.ins[0])):
                if a != b and absorbable(a):
                    st = open_steps[a]
                    if st.res is None and values[a].shape == values[b].shape:
                        pos = steps.index(st)
                        # the residual operand must exist before the conv runs: move the conv to the end unless
                        # something in between overwrites one of its inputs in place
                        later = steps[pos + 1:]
                        if any(s.inplace and s.out is not None and _root(values, s.out) in
                               [_root(values, r) for r in st.reads()] for s in later):
                            continue
                        steps.pop(pos)
                        steps.append(st)
                        st.res, st.out = b, out
                        st.fused.append(nd.name)
                        del open_steps[a]
                        if st.act != 0:
                            st.res_after = True          # Darknet shortcut: x + act(bn(conv)); epilogue is now full
                        else:
                            open_steps[out] = st
                        done = True
                        break
            if done:
                continue
        if k in ACTS and absorbable(x):
            st = open_steps[x]
            if st.act == 0:
                st.act, st.alpha = ACTS[k], float(nd.attrs.get('alpha', 0.2)) if k == 'leakyrelu' else 0.0
                st.out = out
                st.fused.append(nd.name)
                del open_steps[x]
                open_steps[out] = st
                if k == 'relu':
                    values[out].alias_of = None     # the fused value is a fresh buffer, not an alias
                continue
        # ---- standalone launches ----
        if k == 'relu':
            st = emit(Step('relu', nd.name, [x], out))
            st.inplace = True
        elif k == 'identity':
            st = emit(Step('alias', nd.name, [x], out))
        elif k == 'leakyrelu':
            emit(Step('leakyrelu', nd.name, [x], out, {'alpha': float(nd.attrs.get('alpha', 0.2))}))
        elif k == 'sigmoid':
            emit(Step('sigmoid', nd.name, [x], out))
        elif k == 'add':
            emit(Step('add', nd.name, [nd.ins[0], nd.ins[1]], out))
        elif k == 'batchnorm':
            st = emit(Step('scale_shift', nd.name, [x], out))
            st.bn = (nd.ins[1], nd.ins[2])
        elif k == 'clip':
            st = emit(Step('clip', nd.name, [x], out, nd.attrs))       # bounds given as inputs were resolved by infer()
            st.inplace = True
        elif k in ('maxpool', 'averagepool', 'zero_stuff', 'hardsigmoid', 'softmax'):
            emit(Step(k, nd.name, [x], out, nd.attrs))
        elif k in ('upsample', 'resize'):
            emit(Step('upsample', nd.name, [x], out, nd.attrs))
        elif k == 'concat':
            emit(Step('concat', nd.name, list(nd.ins), out))
        elif k == 'gap':
            emit(Step('gap', nd.name, [x], out))
        elif k == 'flatten':
            emit(Step('flatten', nd.name, [x], out))
        else:                                                       # pragma: no cover
            raise NotImplementedError(k)
    return steps


def absorb_shortcuts(plan, eligible):
    """Down-sampling residual blocks: ``conv2 -> bn -> add(., bn_d(conv1x1_d(x))) -> relu``.  When ``eligible(main, short)``
    agrees, the 1x1 shortcut convolution ``short`` (no activation, no residual, result used only as ``main``'s residual)
    is removed from the step list and ``main`` reads the block input itself (``main.shortcut``): one launch instead of
    two, the shortcut's k-chunks accumulate into the same tile (plnr_conv2d_shortcut_fwd)."""
    values, steps = plan.values, plan.steps
    producer = {st.out: st for st in steps}
    n_reads = {}
    for st in steps:
        for r in st.reads():
            n_reads[r] = n_reads.get(r, 0) + 1
    drop = set()
    for st in steps:
        if st.op != 'conv' or st.res is None or st.res_after or st.shortcut:
            continue
        sh = producer.get(st.res)
        if sh is None or sh.op != 'conv' or sh.res is not None or sh.act != 0 or n_reads.get(sh.out, 0) != 1:
            continue
        if values[sh.out].is_output or id(sh) in drop:
            continue
        a = sh.attrs
        if tuple(values[sh.w].shape[2:]) != (1, 1) or a['group'] != 1 or tuple(a['pads']) != (0, 0, 0, 0):
            continue
        if a['strides'][0] != a['strides'][1] or tuple(a['dilations']) != (1, 1):
            continue
        if not eligible(st, sh):
            continue
        st.shortcut = (sh.ins[0], int(a['strides'][0]), sh)
        st.res = None
        st.fused = list(st.fused) + ['(' + '+'.join(sh.fused) + ')']
        drop.add(id(sh))
    if drop:
        plan.steps = [st for st in steps if id(st) not in drop]
    return len(drop)


def _root(values, vid):
    while values[vid].alias_of is not None:
        vid = values[vid].alias_of
    return vid


def _storage_root(values, vid):
    """The value whose buffer holds ``vid``: aliases, then channel slices of a concat output (zero-copy concat)."""
    vid = _root(values, vid)
    while values[vid].slice_of is not None:
        vid = _root(values, values[vid].slice_of[0])
    return vid


def place_concat_inputs(plan, writes_strided=('conv', 'upsample'), reads_strided=('conv', 'concat'), align=8):
    """Zero-copy concat (np.concatenate(axis=1), planer/layer.py:90-91): an input of a channel concat is PRODUCED in place,
    as a channel slice of the concat's output buffer, when its producer can write a strided view (conv epilogues, nearest
    upsample), every reader can read one (convs -- TMA takes any row pitch -- and the concat itself), its channel count and
    offset keep 16-byte alignment, and it is neither a graph output nor already placed elsewhere.  The concat step then
    copies only the inputs that could not be placed.  Returns the number of placed inputs."""
    values, steps = plan.values, plan.steps
    producer = {}
    for st in steps:
        producer.setdefault(_root(values, st.out), st)
    readers = {}
    for st in steps:
        for r in st.reads():
            readers.setdefault(_root(values, r), []).append(st)
    placed = 0
    for st in steps:
        if st.op != 'concat':
            continue
        out = _root(values, st.out)
        if values[out].slice_of is not None:
            continue
        off = 0
        seen = set()
        for i in st.ins:
            c = values[i].shape[1]
            r = _root(values, i)
            v = values[r]
            p = producer.get(r)
            ok = (v.kind == 'act' and not v.is_output and v.slice_of is None and r not in seen and p is not None and
                  p.op in writes_strided and not p.inplace and c % align == 0 and off % align == 0 and
                  all(u.op in reads_strided for u in readers.get(r, [])) and
                  not (p.op == 'upsample' and p.attrs.get('mode', 'nearest') != 'nearest') and
                  not any(values[q].alias_of == r for q in range(len(values))))      # nothing aliases it in place
            if ok and p.op == 'conv' and (p.attrs.get('group', 1) != 1):
                ok = False
            if ok:
                v.slice_of = (out, off)
                placed += 1
            seen.add(r)
            off += c
    return placed


def assign_buffers(plan, elem_bytes, storage_channels):
    """Liveness-based buffer reuse.  Values that alias (in-place relu, identity, flatten of 1x1 maps) share a
    buffer; graph inputs/outputs and weights are never recycled.  ``storage_channels(vid)`` gives the stored
    channel count (inputs may be channel-padded)."""
    values, steps = plan.values, plan.steps
    root = lambda v: _storage_root(values, v)
    for st in steps:                                   # aliases created by standalone steps
        if st.op in ('relu', 'clip', 'alias') or (st.op == 'flatten' and values[st.ins[0]].shape[2:] == (1, 1)):
            values[st.out].alias_of = st.ins[0]
    last_use, size = {}, {}
    for pos, st in enumerate(steps):
        for r in st.reads():
            last_use[root(r)] = pos
        last_use.setdefault(root(st.out), pos)
        last_use[root(st.out)] = max(last_use[root(st.out)], pos)
    pinned = {root(o) for o in plan.outputs} | {root(i) for i in plan.inputs}

    def nbytes(vid):
        shp = values[vid].shape
        n = int(np.prod(shp)) if len(shp) != 4 else shp[0] * storage_channels(vid) * shp[2] * shp[3]
        return max(n, 1) * elem_bytes

    free, buf_of, buf_bytes = {}, {}, []
    for pos, st in enumerate(steps):
        r = root(st.out)
        if r not in buf_of and values[r].kind == 'act':
            need = nbytes(r)
            pool = free.get(need)
            if pool and r not in pinned:
                buf_of[r] = pool.pop()
            else:
                buf_of[r] = len(buf_bytes)
                buf_bytes.append(need)
        for q in {root(x) for x in st.reads()} | {r}:
            if last_use.get(q) == pos and q in buf_of and q not in pinned and values[q].kind == 'act':
                free.setdefault(buf_bytes[buf_of[q]], []).append(buf_of[q])
    plan.buffer_of, plan.buffer_bytes = buf_of, buf_bytes
    return plan


def compile_graph(model, input_shapes, host_consts=None):
    """IR + input shapes -> GraphPlan (no device access)."""
    plan = GraphPlan()
    values, nodes, outputs = flatten(model, input_shapes)
    infer(values, nodes, host_consts)
    plan.values, plan.nodes = values, nodes
    plan.outputs = list(outputs)
    plan.inputs = [v.id for v in values if v.kind == 'input']
    plan.steps = fuse(values, nodes, outputs)
    plan.flops = sum(n.flops for n in nodes)
    return plan
