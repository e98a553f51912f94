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
    return net


def from_model(model, blob, half=False):
    """Build a Net from an in-memory (IR dict, uint8 blob) pair (zoo builders, tests, bench)."""
    net = Net()
    net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
    net.load_weights(blob)
    if half:
        net.half()
    return net
