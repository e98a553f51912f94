"""ONNX -> Planer IR, the reference's ``read_onnx`` / ``onnx2pla`` (planer/io.py:36-299) without the ``onnx`` package.

The reference walks ``onnx.load(path).graph`` (io.py:54-55); this module reads the same protobuf messages with a ~60-line
wire-format decoder (ModelProto.graph = 7; GraphProto.node = 1, initializer = 5, input = 11, output = 12; NodeProto.input = 1,
output = 2, name = 3, op_type = 4, attribute = 5; AttributeProto.name = 1, f = 2, i = 3, s = 4, t = 5, floats = 7, ints = 8;
TensorProto.dims = 1, data_type = 2, float_data = 4, int32_data = 5, int64_data = 7, name = 8, raw_data = 9, double_data = 10)
and emits the IR the reference would: one layer per node, ``flow`` entries ``[inputs, [node name], outputs]``
with one-element lists collapsed to a string (io.py:70-73), every initializer in ``inits`` and in the flat uint8 blob in file
order (io.py:57-63, 286), BatchNormalization folded into ``<gamma>_invK`` / ``<gamma>_invB`` of shape (1, C, 1, 1) with eps
hard-coded to 1e-5 and the four original tensors left in the blob (io.py:76-91), Constant nodes hoisted into ``inits``
(io.py:156-165), a trailing ``return`` layer (io.py:284-285).

Operators: the ones this package implements (``planer_b200.layer_map``); anything else raises NotImplementedError naming the
operator (the reference returns ``('lost', node)``, io.py:281-282).  Deliberate differences, each because the reference's
output cannot run: absent ``strides`` / ``dilations`` / ``pads`` get their ONNX defaults instead of ``None``
(io.py:94-100 would pass ``None`` into ``Conv2d``); ``Clip`` bounds given as inputs (opset >= 11) become attributes (the
reference reads attributes only, io.py:272-278); ``AveragePool`` with ``count_include_pad = 0`` and non-zero pads is
refused (planer/util.py:97-100 always divides by kh*kw).

Validated against files written by PyTorch's own exporter (oracle/gen_onnx_fixtures.py; tests/golden/onnx/).
"""
import struct

import numpy as np

_DTYPES = [None, 'float32', 'uint8', 'int8', 'uint16', 'int16', 'int32', 'int64', 'str', 'bool', 'float16', 'float64',
           'uint32', 'uint64', 'complex64', 'complex128']            # TensorProto.DataType, same table as io.py:36-37


# ---------------------------------------------------------------------------------------------
# protobuf wire format
# ---------------------------------------------------------------------------------------------
def _varint(buf, pos):
    val, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    """Yield (field number, wire type, value) of one message; length-delimited values are memoryview slices."""
    buf = memoryview(buf)
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + n], pos + n
        elif wt == 5:
            val, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield num, wt, val


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(wt, val, out):
    if wt == 0:
        out.append(_signed(val))
    else:
        pos, end = 0, len(val)
        while pos < end:
            v, pos = _varint(val, pos)
            out.append(_signed(v))


def _tensor(buf):
    """TensorProto -> (name, numpy array) like ``onnx.numpy_helper.to_array``."""
    dims, dtype, name, raw = [], 1, '', None
    floats, i32, i64, f64 = [], [], [], []
    for num, wt, val in _fields(buf):
        if num == 1: _packed_varints(wt, val, dims)
        elif num == 2: dtype = val
        elif num == 8: name = bytes(val).decode()
        elif num == 9: raw = bytes(val)
        elif num == 4:
            floats.extend(struct.unpack('<%df' % (len(val) // 4), bytes(val)) if wt == 2 else struct.unpack('<f', val))
        elif num == 5: _packed_varints(wt, val, i32)
        elif num == 7: _packed_varints(wt, val, i64)
        elif num == 10:
            f64.extend(struct.unpack('<%dd' % (len(val) // 8), bytes(val)) if wt == 2 else struct.unpack('<d', val))
        elif num == 13:
            raise NotImplementedError('tensor %r keeps its data in an external file' % name)
    np_dtype = _DTYPES[dtype] if 0 < dtype < len(_DTYPES) else None
    if np_dtype in (None, 'str'):
        raise NotImplementedError('tensor %r has unsupported ONNX data type %d' % (name, dtype))
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dtype).copy()
    elif np_dtype == 'float16':
        arr = np.array(i32, dtype=np.uint16).view(np.float16)          # fp16 bit patterns travel in int32_data
    elif floats: arr = np.array(floats, dtype=np_dtype)
    elif f64: arr = np.array(f64, dtype=np_dtype)
    elif i64: arr = np.array(i64, dtype=np_dtype)
    else: arr = np.array(i32, dtype=np_dtype)
    return name, arr.reshape(dims)


def _attribute(buf):
    """AttributeProto -> (name, value).  Field 20 (``type``: 1 FLOAT, 2 INT, 3 STRING, 4 TENSOR, 6 FLOATS, 7 INTS) tells a
    scalar whose value is the wire default (0 / 0.0 / '' may be left out by a writer) from an empty list."""
    name, f, i, s, t, floats, ints, kind = '', None, None, None, None, [], [], 0
    for num, wt, val in _fields(buf):
        if num == 20: kind = int(val)
        if num == 1: name = bytes(val).decode()
        elif num == 2: f = struct.unpack('<f', val)[0]
        elif num == 3: i = _signed(val)
        elif num == 4: s = bytes(val).decode()
        elif num == 5: t = _tensor(val)[1]
        elif num == 7: floats.extend(struct.unpack('<%df' % (len(val) // 4), bytes(val)) if wt == 2 else struct.unpack('<f', val))
        elif num == 8: _packed_varints(wt, val, ints)
    for v in (t, s, f, i):
        if v is not None:
            return name, v
    if kind in (1, 2, 3):
        return name, {1: 0.0, 2: 0, 3: ''}[kind]
    return name, (ints if ints or kind == 7 else floats)


def _node(buf):
    nd = {'input': [], 'output': [], 'name': '', 'op_type': '', 'attrs': {}}
    for num, wt, val in _fields(buf):
        if num == 1: nd['input'].append(bytes(val).decode())
        elif num == 2: nd['output'].append(bytes(val).decode())
        elif num == 3: nd['name'] = bytes(val).decode()
        elif num == 4: nd['op_type'] = bytes(val).decode()
        elif num == 5:
            k, v = _attribute(val)
            nd['attrs'][k] = v
    return nd


def _value_name(buf):
    for num, wt, val in _fields(buf):
        if num == 1:
            return bytes(val).decode()
    return ''


def parse_graph(data):
    """ModelProto bytes -> {'nodes', 'initializers' [(name, array)], 'inputs', 'outputs'} in file order."""
    graph = None
    for num, wt, val in _fields(data):
        if num == 7:
            graph = val
    if graph is None:
        raise ValueError('no GraphProto in the ONNX model')
    g = {'nodes': [], 'initializers': [], 'inputs': [], 'outputs': []}
    for num, wt, val in _fields(graph):
        if num == 1: g['nodes'].append(_node(val))
        elif num == 5: g['initializers'].append(_tensor(val))
        elif num == 11: g['inputs'].append(_value_name(val))
        elif num == 12: g['outputs'].append(_value_name(val))
    return g


# ---------------------------------------------------------------------------------------------
# graph -> IR
# ---------------------------------------------------------------------------------------------
def _ints(a, key, default):
    v = a.get(key)
    return list(default) if v is None or v == [] else [int(i) for i in v]


def read_onnx(path_or_bytes):
    """-> (model dict {'input', 'inits', 'layers', 'flow'}, flat uint8 weight blob), planer/io.py:53-287."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, 'rb').read()
    g = parse_graph(data)
    inits, weights, where = [], [], {}

    def add_init(name, v):
        where[name] = len(weights)
        inits.append([name, list(v.shape), str(v.dtype)])
        weights.append(v.reshape(1) if v.ndim == 0 else v)             # a scalar occupies one element (io.py:62,164)

    for name, v in g['initializers']:
        add_init(name, v)
    const = lambda name: weights[where[name]] if name in where else None
    layers, flows = [], []
    for n_i, nd in enumerate(g['nodes']):
        op, a = nd['op_type'], nd['attrs']
        lname = nd['name'] or '%s_%d' % (op, n_i)
        ins, outs = list(nd['input']), list(nd['output'])
        flow = [ins[0] if len(ins) == 1 else ins, [lname], outs[0] if len(outs) == 1 else outs]
        if op == 'Constant':                                           # io.py:156-165: hoisted, no layer, no flow
            v = a.get('value')
            if v is None:
                raise NotImplementedError('Constant node %r without a tensor value' % lname)
            add_init(outs[0], np.asarray(v))
            continue
        if op == 'BatchNormalization':                                 # io.py:76-91
            k, b, m, v = [const(ins[j]) for j in (1, 2, 3, 4)]
            if any(t is None for t in (k, b, m, v)):
                raise NotImplementedError('BatchNormalization %r with non-constant statistics' % lname)
            v_inv = 1 / np.sqrt(v + 1e-5)
            inv_b, inv_k = (-k * m * v_inv + b).reshape(1, -1, 1, 1), (k * v_inv).reshape(1, -1, 1, 1)
            kname, bname = ins[1] + '_invK', ins[1] + '_invB'
            add_init(kname, inv_k)
            add_init(bname, inv_b)
            flow[0] = [ins[0], kname, bname]
            layers.append([lname, 'batchnorm', {}])
        elif op == 'Conv':
            layers.append([lname, 'conv', {'group': int(a.get('group') or 1), 'strides': _ints(a, 'strides', (1, 1)),
                                           'dilations': _ints(a, 'dilations', (1, 1)), 'pads': _ints(a, 'pads', (0, 0, 0, 0))}])
        elif op == 'ConvTranspose':
            layers.append([lname, 'convtranspose', {'group': int(a.get('group') or 1), 'strides': _ints(a, 'strides', (1, 1)),
                                                    'dilations': _ints(a, 'dilations', (1, 1)),
                                                    'pads': _ints(a, 'pads', (0, 0, 0, 0)),
                                                    'output_padding': _ints(a, 'output_padding', (0, 0))}])
        elif op == 'Gemm':
            w = const(ins[1])
            if w is None or int(a.get('transB', 0)) != 1 or int(a.get('transA', 0)) != 0 or \
                    float(a.get('alpha', 1.0)) != 1.0 or float(a.get('beta', 1.0)) != 1.0:
                raise NotImplementedError('Gemm %r: only y = x @ W.T + b with a constant W (torch.nn.Linear) maps to the '
                                          'reference\'s dense layer (planer/layer.py:15-18)' % lname)
            if len(ins) == 2:
                raise NotImplementedError('Gemm %r without bias: the reference\'s Dense needs one (planer/layer.py:15)' % lname)
            layers.append([lname, 'dense', {'shp': list(w.shape[::-1])}])
        elif op in ('MaxPool', 'AveragePool'):
            w = _ints(a, 'kernel_shape', ())
            pads = _ints(a, 'pads', (0,) * (2 * len(w)))
            if int(a.get('ceil_mode', 0)) != 0:
                raise NotImplementedError('%s %r: ceil_mode (the reference floors, planer/util.py:84-85)' % (op, lname))
            if op == 'AveragePool' and not int(a.get('count_include_pad', 0)) and any(pads):
                raise NotImplementedError('AveragePool %r: count_include_pad = 0 with padding (the reference always divides '
                                          'by kh*kw, planer/util.py:97-100)' % lname)
            layers.append([lname, 'maxpool' if op == 'MaxPool' else 'averagepool',
                           {'w': w, 'pads': pads, 'strides': _ints(a, 'strides', (1,) * len(w))}])
        elif op == 'GlobalAveragePool': layers.append([lname, 'gap', {}])
        elif op == 'Upsample': layers.append([lname, 'upsample', {'mode': a.get('mode', 'nearest')}])
        elif op == 'Resize':
            layers.append([lname, 'resize', {'mode': a.get('mode', 'nearest'),
                                             'nearest_mode': a.get('nearest_mode', 'round_prefer_floor'),
                                             'coordinate_transformation_mode': a.get('coordinate_transformation_mode', 'half_pixel')}])
        elif op == 'Flatten':
            if int(a.get('axis', 1)) != 1:
                raise NotImplementedError('Flatten %r with axis != 1 (planer/layer.py:59 keeps the batch axis)' % lname)
            layers.append([lname, 'flatten', {}])
        elif op == 'Relu': layers.append([lname, 'relu', {}])
        elif op == 'LeakyRelu': layers.append([lname, 'leakyrelu', {'alpha': float(a.get('alpha', 0.01))}])
        elif op == 'HardSigmoid':
            layers.append([lname, 'hardsigmoid', {k: float(a[k]) for k in ('alpha', 'beta') if k in a}])
        elif op == 'Sigmoid': layers.append([lname, 'sigmoid', {}])
        elif op == 'Add': layers.append([lname, 'add', {}])
        elif op == 'Identity': layers.append([lname, 'identity', {}])
        elif op == 'Concat': layers.append([lname, 'concat', {'axis': int(a.get('axis', 0))}])
        elif op == 'Softmax': layers.append([lname, 'softmax', {'axis': int(a.get('axis', -1))}])
        elif op == 'Clip':
            para = {}
            for key, idx in (('min', 1), ('max', 2)):
                if key in a:
                    para[key] = float(a[key])
                elif len(ins) > idx and ins[idx]:
                    c = const(ins[idx])
                    if c is None:
                        raise NotImplementedError('Clip %r with a computed %s bound' % (lname, key))
                    para[key] = float(np.asarray(c).reshape(-1)[0])
            if 'min' not in para or 'max' not in para:
                raise NotImplementedError('Clip %r with an open bound (the reference\'s defaults are min=0, max=1, '
                                          'planer/layer.py:247)' % lname)
            flow[0] = ins[0]
            layers.append([lname, 'clip', para])
        else:
            raise NotImplementedError('ONNX operator %r (node %r) is not implemented by planer_b200; the reference reports it '
                                      'as a lost layer (planer/io.py:281-282)' % (op, lname))
        flows.append(flow)
    layers.append(['return', 'return', {}])
    flows.append([list(g['outputs']), ['return'], 'plrst'])
    blob = np.concatenate([np.ascontiguousarray(w).reshape(-1).view(np.uint8) for w in weights]) if weights else np.zeros(0, np.uint8)
    return {'input': list(g['inputs']), 'inits': inits, 'layers': layers, 'flow': flows}, blob


def onnx2pla(path, zip=True):
    """planer/io.py:289-299: write ``<path without .onnx>.pla`` (or .json + .npy) next to the ONNX file."""
    from . import zoo
    model, blob = read_onnx(path)
    zoo.save_model(path[:-5] if path.endswith('.onnx') else path, model, blob, pla=zip)
    return model, blob
