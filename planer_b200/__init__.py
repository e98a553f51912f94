"""planer_b200 -- the B200-native (sm_100a) forward hot path of Image-Py/planer behind Planer's own API.

    import planer_b200 as planer
    planer.core(planer.b200)                  # the backend hook of planer/__init__.py:22-38
    net = planer.read_net('resnet18')         # .json + .npy / .pla, planer/io.py:8-34
    net.half()                                # planer/net.py:26-29
    y = net(x)                                # numpy NCHW in, numpy out, planer/net.py:94-101
    for y in net.map(batches): ...            # the same call, pipelined over a stream of host batches (net.py)

``core(obj)`` accepts the B200 backend module (``planer_b200.b200``); there is NO CPU path in this package:
passing numpy raises.  ``install(planer)`` plugs the same kernels into the reference package's operator
table so the reference ``Net`` drives them.

Importing this package never touches the GPU, creates no directories and prints nothing (the reference does
all three at import: planer/__init__.py:19-20,50-51).
"""
from . import backend as b200
from .layer import wrap, layer_map
from .net import Net
from .io import read_net, from_model, save_pack, load_pack
from .onnx_import import read_onnx, onnx2pla
from . import zoo
from . import util
from .util import tile, resize

InferenceSession = read_net          # planer/__init__.py:7

__version__ = '0.1.0'
backend = b200


def core(obj=None, silent=True):
    """planer/__init__.py:22-38: select the array backend.  Only the B200 backend exists here."""
    if obj is None or obj is b200:
        if not silent: print('\nuser switch engine:', b200.__name__)
        return b200
    name = getattr(obj, '__name__', repr(obj))
    raise NotImplementedError(
        'planer_b200.core(%s): this package implements the B200 (sm_100a) path only and has no CPU or '
        'multi-backend fallback; pass planer_b200.b200 (use the reference package for numpy/cupy)' % name)


def asnumpy(arr, **key): return b200.asnumpy(arr)


def asarray(arr, **key): return b200.asarray(arr, **key)


def pinned_empty(shape, dtype='float32'): return b200.pinned_empty(shape, dtype)


def install(planer):
    """Drop the B200 operators into the REFERENCE package ``planer`` (Level B of SURVEY 8b): its own
    ``Net`` / ``read_net`` then run on the GPU.  Equivalent to the two lines a maintainer would add to
    planer/__init__.py."""
    planer.core(b200, True)                       # rebinds np in util/layer/net/io (planer/__init__.py:23-25)
    for kind, fn in layer_map.items():
        planer.layer.layer_map[kind] = fn
    return planer
