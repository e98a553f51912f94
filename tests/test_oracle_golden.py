"""Pin the oracle (oracle/planer_oracle.py) against outputs of the unmodified reference.

Fixtures under tests/golden/ were produced by oracle/gen_golden.py, which imports
/root/reference/planer and runs its numpy path.  The oracle restates the same numpy calls in the
same order, so every comparison here is BIT-EXACT (integer compare of the raw bytes), fp32 and
fp16 alike.  The slow true-fp16 ResNet-18 case (numpy fp16 matmul has no BLAS) is included once.
"""
import hashlib
import os

import numpy as np
import pytest

import planer_oracle as oracle
from tests import cases

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
OPS = np.load(os.path.join(GOLD, 'ops.npz'))
GRAPHS = np.load(os.path.join(GOLD, 'graphs.npz'))


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    # -0.0 vs +0.0 (quirk Q5) would differ bitwise only if the op order differed; we demand equality.
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize('name', list(cases.OP_CASES))
def test_op_bit_exact(name):
    kind, args, kw = cases.make_case(name)
    raw = np.concatenate([np.ascontiguousarray(a).reshape(-1).view(np.uint8) for a in args if isinstance(a, np.ndarray)])
    assert _digest(raw) == str(OPS[name + '.sha']), 'seeded inputs drifted from the fixture'
    y = oracle.layer_map[kind](*[a.copy() for a in args], **kw)
    _same_bits(np.ascontiguousarray(y), OPS[name])


@pytest.mark.parametrize('name', list(cases.GRAPH_CASES))
def test_graph_bit_exact(name):
    model, blob, x, half = cases.make_graph_case(name)
    assert _digest(x) == str(GRAPHS[name + '.sha_x'])
    assert _digest(blob) == str(GRAPHS[name + '.sha_blob'])
    net = oracle.build_net(model, blob, half)
    y = net(x.copy())
    ys = y if isinstance(y, tuple) else (y,)
    assert len(ys) == int(GRAPHS[name + '.nout'])
    for i, t in enumerate(ys):
        assert tuple(GRAPHS['%s.shape%d' % (name, i)]) == t.shape
        _same_bits(cases.sample(t), GRAPHS['%s.out%d' % (name, i)])


def test_maxpool_quirk_zero_pad_and_floor():
    """padding value is 0 (not -inf) and the accumulator floor is -1e4."""
    x = np.full((1, 1, 4, 4), -5.0, np.float32)
    y = oracle.maxpool(x, (3, 3), (1, 1, 1, 1), (2, 2))
    assert y[0, 0, 0, 0] == 0.0          # window touches the zero padding
    assert y[0, 0, 1, 1] == -5.0         # interior window
    assert oracle.maxpool(np.full((1, 1, 2, 2), -3e4, np.float32))[0, 0, 0, 0] == np.float32(-1e4)


def test_relu_is_in_place_and_aliases():
    """ReLU mutates and returns its input object."""
    x = np.array([-1.0, 2.0], np.float32)
    assert oracle.relu(x) is x and x[1] == 2.0 and x[0] == 0.0


def test_batchnorm_fold_matches_formula():
    rng = np.random.default_rng(3)
    g, b, m = (rng.standard_normal(8).astype(np.float32) for _ in range(3))
    v = rng.uniform(0.5, 1.5, 8).astype(np.float32)
    k, s = oracle.fold_batchnorm(g, b, m, v)
    x = rng.standard_normal((2, 8, 3, 3)).astype(np.float32)
    ref = g.reshape(1, -1, 1, 1) * (x - m.reshape(1, -1, 1, 1)) / np.sqrt(v.reshape(1, -1, 1, 1) + 1e-5) + b.reshape(1, -1, 1, 1)
    assert np.allclose(oracle.batchnorm(x, k, s), ref, atol=1e-5)


def _reference_module():
    """The UNMODIFIED reference, pip-installed into baseline/_ref by __graft_entry__.build() (numpy backend)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_root = os.path.join(root, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'planer')):
        pytest.skip('baseline/_ref is not in this checkout')
    if not os.access(os.path.expanduser('~'), os.W_OK):
        os.environ['HOME'] = '/tmp'
    sys.path.insert(0, ref_root)
    try:
        import planer as ref
    finally:
        sys.path.remove(ref_root)
    import numpy
    ref.core(numpy, True)
    return ref


@pytest.mark.parametrize('half,count,batch', [(False, 24, 2), (True, 8, 1)])
def test_oracle_equals_the_reference_itself_on_random_graphs(half, count, batch):
    """The oracle's interpreter against the UNMODIFIED reference's Net (baseline/_ref, numpy backend) on random DAGs of the
    hot-path operators (the generator of tests/test_gpu_parity.py): BIT-exact, like on the committed fixtures -- in float32
    and after ``Net.half()`` (numpy's float16 matmul has no BLAS: fewer, single-image graphs)."""
    ref = _reference_module()
    for seed in range(count):
        model, blob, cin, size = cases.random_graph((9100 if half else 9000) + seed)
        x = np.random.default_rng(seed).standard_normal((batch, cin) + tuple(size)).astype(np.float16 if half else np.float32)
        net = ref.Net()
        net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
        net.load_weights(blob)
        if half:
            net.half()
        if hasattr(ref.util, 'clear_buf'):
            ref.util.clear_buf()
        want = net(x.copy())
        want = want if isinstance(want, tuple) else (want,)
        got = oracle.build_net(model, blob, half=half)(x.copy())
        got = got if isinstance(got, tuple) else (got,)
        assert len(got) == len(want)
        for g, w_ in zip(got, want):
            g, w_ = np.asarray(g), np.asarray(w_)
            assert g.dtype == w_.dtype and np.array_equal(g, w_), (seed, float(np.abs(g.astype(np.float32) - w_.astype(np.float32)).max()))
