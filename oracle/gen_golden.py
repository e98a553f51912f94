"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/planer).

Run once in the authoring container (the reference mount does not exist on the GPU box):

    HOME=/tmp/planer_home PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Inputs and weights are regenerated from seeds by the tests, so fixtures hold only the reference
outputs plus sha256 digests of the seeded inputs/blobs (to catch RNG drift).  Large outputs are
stored as a strided sample.  TEST INFRASTRUCTURE ONLY -- nothing in planer_b200/ imports this.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
os.environ.setdefault('HOME', '/tmp/planer_home')
os.makedirs(os.environ['HOME'], exist_ok=True)

import planer                      # noqa: E402  (the reference)
from planer import layer as L      # noqa: E402
from planer_b200 import zoo        # noqa: E402  (IR builders only; no GPU code is touched)

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_net(model, blob, half=False):
    net = planer.Net()
    net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
    net.load_weights(blob)
    if half:
        net.half()
    return net


def op_cases():
    """Seeded per-op cases; shared with tests/cases.py via the same generator function."""
    from tests.cases import OP_CASES, make_case
    path = os.path.join(OUT, 'ops.npz')
    out = dict(np.load(path)) if INCREMENTAL and os.path.exists(path) else {}
    for name in OP_CASES:
        if name in out:
            continue
        kind, args, kw = make_case(name)
        fn = L.layer_map[kind]
        planer.util.clear_buf()   # quirk Q6: the global im2col scratch keeps its dtype between calls
        y = fn(*[a.copy() if isinstance(a, np.ndarray) else a for a in args], **kw)
        out[name] = np.ascontiguousarray(y)
        out[name + '.sha'] = np.array(digest(np.concatenate(
            [np.ascontiguousarray(a).reshape(-1).view(np.uint8) for a in args if isinstance(a, np.ndarray)])))
    np.savez_compressed(os.path.join(OUT, 'ops.npz'), **out)
    print('ops.npz', len(OP_CASES), 'cases')


def graph_cases():
    from tests.cases import GRAPH_CASES, make_graph_case, sample
    path = os.path.join(OUT, 'graphs.npz')
    out = dict(np.load(path)) if INCREMENTAL and os.path.exists(path) else {}
    for name in GRAPH_CASES:
        if name + '.nout' in out:
            continue
        model, blob, x, half = make_graph_case(name)
        planer.util.clear_buf()   # quirk Q6: a scratch buffer left by an earlier fp16 conv would make this graph's im2col fp16
        net = ref_net(model, blob, half)
        y = net(x.copy())
        ys = y if isinstance(y, tuple) else (y,)
        for i, t in enumerate(ys):
            t = np.ascontiguousarray(t)
            out['%s.out%d' % (name, i)] = sample(t)
            out['%s.shape%d' % (name, i)] = np.array(t.shape)
            out['%s.absmax%d' % (name, i)] = np.array(np.abs(t.astype(np.float64)).max())
        out[name + '.nout'] = np.array(len(ys))
        out[name + '.sha_x'] = np.array(digest(x))
        out[name + '.sha_blob'] = np.array(digest(blob))
        print(name, [t.shape for t in ys], flush=True)
    np.savez_compressed(os.path.join(OUT, 'graphs.npz'), **out)


def tile_cases():
    """The reference's sliding-window decorator (planer/util.py:291-348) around deterministic per-window functions."""
    from tests.cases import TILE_CASES, make_tile_case
    path = os.path.join(OUT, 'tile.npz')
    if INCREMENTAL and os.path.exists(path) and all(n in np.load(path) for n in TILE_CASES):
        return
    out = {}
    for name in TILE_CASES:
        img, kw, fn = make_tile_case(name)
        y = planer.util.tile(progress=lambda *a: None, **kw)(fn)(img.copy())
        out[name] = np.ascontiguousarray(y)
        out[name + '.sha'] = np.array(digest(img))
    np.savez_compressed(os.path.join(OUT, 'tile.npz'), **out)
    print('tile.npz', len(TILE_CASES), 'cases')


# ``--add``: keep the fixtures already in the .npz files and run the reference only for cases that are not there yet
INCREMENTAL = '--add' in sys.argv

if __name__ == '__main__':
    tile_cases()
    op_cases()
    graph_cases()
