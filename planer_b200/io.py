"""Model reader with the reference's ``read_net`` contract (planer/io.py:8-34).

Formats: ``<name>.pla`` (zip of ``<name>.json`` + ``<name>.npy``, io.py:12-18,295-297) or ``<name>.json`` +
``<name>.npy`` (io.py:19-24); the ``.npy`` is the flat uint8 concatenation of every init (io.py:286).  Files
written by the reference load here byte for byte and vice versa (``zoo.save_model``).

``<name>.onnx`` goes through ``onnx_import.read_onnx`` (io.py:36-299 restated without the ``onnx`` package, SURVEY 8f
rank 1): graphs made of the implemented operators load, anything else raises NotImplementedError naming the operator.
"""
import json
import os
import zipfile
from io import BytesIO

import numpy

from . import dist
from .net import Net


def read_net(path, debug=False):
    net = Net()
    path = path.replace('.onnx', '')
    rank, world = dist.rank_world()
    reader = rank == 0 or world == 1          # only rank 0 touches the .npy; the others get the broadcast
    weights = None
    if os.path.exists(path + '.pla'):
        with zipfile.ZipFile(path + '.pla') as f:
            base = os.path.split(path)[1]
            body = json.loads(f.read(base + '.json'))
            if reader:
                weights = numpy.load(BytesIO(f.read(base + '.npy')))
    elif os.path.exists(path + '.json'):
        with open(path + '.json') as f:
            body = json.load(f)
        if reader:
            weights = numpy.load(path + '.npy')
    elif os.path.exists(path + '.onnx'):
        from .onnx_import import read_onnx                    # planer/io.py:25-29; unsupported operators raise by name
        body, weights = read_onnx(path + '.onnx')
        if not reader:
            weights = None
    else:
        return print('model %s not found!' % path)           # planer/io.py:30-31
    net.load_json(body['input'], body['inits'], body['layers'], body['flow'], debug)
    net.load_weights(weights)
    net._pack_path = path + PACK_SUFFIX           # pre-packed weight cache beside the model (save_pack / load at first use)
    return net


# ---------------------------------------------------------------------------------------------------------------------
# Pre-packed weight cache.  The reference re-derives nothing at load (its weights are used as stored,
# planer/net.py:83-88; Net.half() casts per call site, planer/net.py:26-29); here every executor build used to re-run the
# one-off device work -- fp32 -> fp16 casts, OIHW -> [Cout][kh][kw][Cin] packing, BatchNorm / bias folding, first-layer and
# shortcut packings: ~94 launches for ResNet-18.  `save_pack(net)` writes what the net's pack store holds to
# `<model>.b200pack.npz`, keyed by the SHA-256 of the weight blob, the ABI version and the library's pack-format version;
# a later `read_net` + first forward uploads the cached arrays instead (stale or foreign caches are ignored, never trusted).
# ---------------------------------------------------------------------------------------------------------------------
PACK_SUFFIX = '.b200pack.npz'
PACK_FORMAT = 1


def _blob_digest(net):
    import hashlib
    from . import backend as B
    n = sum(int(numpy.prod(s)) * d.itemsize for _, s, d in net._init_meta)
    flat = B.DeviceArray(net._blob.buf, (n,), numpy.uint8, 'flat', offset=net._blob.offset)
    return hashlib.sha256(flat.get().tobytes()).hexdigest()


def save_pack(net, path=None):
    """Write the one-off weight artefacts the net has built so far (run at least one forward per input signature you
    want covered) to ``path`` (default: beside the model file ``read_net`` loaded).  Returns the path."""
    from . import backend as B, _capi
    path = path or getattr(net, '_pack_path', None)
    if path is None:
        raise ValueError('save_pack: no path given and the net was not loaded from a file')
    B.synchronize()
    out = {'__meta__': numpy.array(json.dumps({'blob_sha256': _blob_digest(net), 'abi': _capi.load().plnr_abi_version(),
                                               'format': PACK_FORMAT}))}
    for k, v in net._pack_store.items():
        out[k] = numpy.array(v) if k.endswith('#n') else v.get()
    tmp = path + '.tmp.npz'
    numpy.savez(tmp, **out)
    os.replace(tmp, path)
    return path


def load_pack(net, path=None):
    """Fill the net's pack store from a cache written by ``save_pack``.  Returns the number of arrays taken (0 when the
    file is missing, was written for other weights / another ABI, or is unreadable)."""
    from . import backend as B, _capi
    path = path or getattr(net, '_pack_path', None)
    if path is None or not os.path.exists(path) or net._blob is None:
        return 0
    try:
        z = numpy.load(path, allow_pickle=False)
        meta = json.loads(str(z['__meta__']))
        if meta.get('format') != PACK_FORMAT or meta.get('abi') != _capi.load().plnr_abi_version() or \
                meta.get('blob_sha256') != _blob_digest(net):
            return 0
        n = 0
        for k in z.files:
            if k == '__meta__':
                continue
            if k.endswith('#n'):
                net._pack_store[k] = int(z[k])
            else:
                net._pack_store[k] = B.asarray(z[k])
                n += 1
        return n
    except Exception:
        net._pack_store = {}
        return 0


def from_model(model, blob, half=False):
    """Build a Net from an in-memory (IR dict, uint8 blob) pair (zoo builders, tests, bench)."""
    net = Net()
    net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
    net.load_weights(blob)
    if half:
        net.half()
    return net
