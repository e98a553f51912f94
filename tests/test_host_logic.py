"""CPU tests of the host side: C-ABI surface, flow compiler (fusion / liveness / shapes / FLOPs), Net plumbing,
model files, multi-process weight broadcast over gloo.  No GPU, no compute through the library."""
import ctypes
import io
import json
import os
import re
import contextlib

import numpy as np
import pytest

import planer_oracle as oracle
import planer_b200 as planer
from planer_b200 import plan as P, zoo, _capi, dist
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    return _capi.load()


def test_library_exports_every_declared_symbol(lib):
    """include/planer_b200.h is the contract: every declared function must be exported and bound."""
    hdr = open(os.path.join(ROOT, 'include', 'planer_b200.h')).read()
    declared = set(re.findall(r'\b(plnr_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 30
    raw = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), 'header declares %s but the library does not export it' % name
    assert declared - {'plnr_last_error'} == set(_capi.PROTOTYPES), 'ctypes binding and header differ'
    assert lib.plnr_abi_version() == 3


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_capi.Tensor) == 32
    assert ctypes.sizeof(_capi.ConvDesc) == 13 * 4
    assert ctypes.sizeof(_capi.Epilogue) == 64 and _capi.Epilogue.acc_scale_dev.offset == 48 and _capi.Epilogue.pool_sum.offset == 56
    assert _capi.Epilogue.act.offset == 24 and _capi.Epilogue.res_after_act.offset == 32 and _capi.Epilogue.out_nchw.offset == 36
    assert _capi.Epilogue.out_f32.offset == 40 and _capi.Epilogue.acc_scale.offset == 44


def test_no_cpu_fallback(lib):
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_capi.PlanerB200Error):
        planer.b200.init()
    with pytest.raises(NotImplementedError):
        planer.core(np)
    with pytest.raises(NotImplementedError):
        planer.layer_map['lstm']
    assert planer.core(planer.b200) is planer.b200


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'planer_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'planer_oracle' not in src and 'import oracle' not in src, fn


# ---------------------------------------------------------------------------------------------
# flow compiler
# ---------------------------------------------------------------------------------------------

def test_resnet18_plan_fuses_to_24_launch_steps():
    model, _ = cases.get_model('resnet18')
    gp = P.compile_graph(model, {'x': (128, 3, 224, 224)})
    assert gp.summary() == {'conv': 20, 'maxpool': 1, 'gap': 1, 'flatten': 1, 'dense': 1}
    # 3.627123 GFLOP conv + 0.001024 dense per image
    assert gp.flops == 128 * 3628146688
    fused = {s.name: s for s in gp.steps}
    assert fused['conv1'].fused == ['conv1', 'bn1', 'relu'] and fused['conv1'].act == 1
    blk = fused['layer1.0.conv2']
    assert blk.fused == ['layer1.0.conv2', 'layer1.0.bn2', 'layer1.0.add', 'layer1.0.relu2'] and blk.res is not None
    # downsample block: the 1x1 conv runs BEFORE conv2, which takes its output as residual
    names = [s.name for s in gp.steps]
    assert names.index('layer2.0.downsample.0') < names.index('layer2.0.conv2')
    assert fused['layer2.0.downsample.0'].act == 0 and fused['layer2.0.conv2'].res is not None
    assert gp.values[gp.outputs[0]].shape == (128, 1000)


def test_yolov3_plan_shapes_and_darknet_shortcut_fusion():
    model, _ = cases.get_model('yolov3_quarter')
    consts = {n: np.array([1, 1, 2, 2], np.float32) for n, s, d in model['inits'] if 'scales' in n}
    gp = P.compile_graph(model, {'x': (2, 3, 96, 96)}, consts)
    assert gp.summary() == {'conv': 75, 'upsample': 2, 'concat': 2}       # 23 shortcut adds live in conv epilogues
    assert sum(1 for s in gp.steps if s.res_after) == 23
    assert [gp.values[o].shape for o in gp.outputs] == [(2, 255, 3, 3), (2, 255, 6, 6), (2, 255, 12, 12)]


def test_yolov3_416_flops_match_survey():
    shp = P.infer_shapes(*[zoo.yolov3(0)[0]], {'x': (1, 3, 416, 416)})
    total = sum(n['flops'] for n in shp['nodes'])
    assert abs(total / 1e9 - 65.864) < 0.01           # Darknet's 65.86 BFLOPs
    assert shp['outputs'] == [(1, 255, 13, 13), (1, 255, 26, 26), (1, 255, 52, 52)]


def test_fusion_respects_multiple_consumers_and_outputs():
    b = zoo._Builder(0)
    y = b.conv('x', 8, 8, 3)
    r = b.op('relu', {}, [y])
    z = b.op('add', {}, [r, y])          # y (== r after the in-place relu) is read twice
    model, _ = b.finish(['x'], [z, r])
    gp = P.compile_graph(model, {'x': (1, 8, 8, 8)})
    # the in-place relu renames y -> r for every later reader (App. D Q4), so the conv output has ONE reader (the
    # relu) and is folded; the add then reads the relu'd value twice, exactly what the reference computes
    assert [s.op for s in gp.steps] == ['conv', 'add'] and gp.steps[0].act == 1
    add = gp.steps[1]
    assert add.ins[0] == add.ins[1] == gp.steps[0].out
    # a second consumer of the PRE-activation value blocks the fold (leakyrelu is not in place)
    b = zoo._Builder(0)
    y = b.conv('x', 8, 8, 3)
    r = b.op('leakyrelu', {'alpha': 0.1}, [y])
    z = b.op('add', {}, [r, y])
    model, _ = b.finish(['x'], [z])
    gp = P.compile_graph(model, {'x': (1, 8, 8, 8)})
    assert [s.op for s in gp.steps] == ['conv', 'leakyrelu', 'add']
    # a graph output is never folded away
    b = zoo._Builder(0)
    y = b.conv('x', 8, 8, 3)
    r = b.op('leakyrelu', {'alpha': 0.1}, [y])
    model, _ = b.finish(['x'], [y, r])
    gp = P.compile_graph(model, {'x': (1, 8, 8, 8)})
    assert [s.op for s in gp.steps] == ['conv', 'leakyrelu']


def test_liveness_reuses_buffers_but_never_outputs():
    model, _ = cases.get_model('resnet18')
    gp = P.compile_graph(model, {'x': (8, 3, 224, 224)})
    P.assign_buffers(gp, 2, lambda v: gp.values[v].shape[1])
    n_values = len({P._root(gp.values, s.out) for s in gp.steps})
    assert len(gp.buffer_bytes) < n_values            # reuse happened
    # no step may write a buffer that one of its own inputs lives in (unless it is the in-place alias)
    for s in gp.steps:
        out_b = gp.buffer_of.get(P._root(gp.values, s.out))
        for r in s.reads():
            rr = P._root(gp.values, r)
            if rr != P._root(gp.values, s.out) and rr in gp.buffer_of:
                assert gp.buffer_of[rr] != out_b, s.name


def test_liveness_simulation_no_overlap():
    """Replay the schedule: a buffer may be rewritten only after the last read of its previous tenant."""
    model, _ = cases.get_model('yolov3_quarter')
    consts = {n: np.array([1, 1, 2, 2], np.float32) for n, s, d in model['inits'] if 'scales' in n}
    gp = P.compile_graph(model, {'x': (1, 3, 64, 64)}, consts)
    P.assign_buffers(gp, 2, lambda v: gp.values[v].shape[1])
    root = lambda v: P._root(gp.values, v)
    last = {}
    for pos, s in enumerate(gp.steps):
        for r in s.reads():
            last[root(r)] = pos
    tenant = {}
    for pos, s in enumerate(gp.steps):
        o = root(s.out)
        if o in gp.buffer_of:
            b = gp.buffer_of[o]
            prev = tenant.get(b)
            if prev is not None and prev != o:
                assert last.get(prev, -1) < pos, (s.name, prev, o)
            tenant[b] = o


def _emulate_stem_contraction(x, W2, kh, kw, pad_t, pad_l, e_min, taps):
    """numpy restatement of what csrc/stem_pool.cu contracts: packed rows t = input rows (2t, 2t+1), operand
    A_t[ow, (ph*3 + c)*8 + j] = x[c, 2t + ph, 2ow + j - P] (P = pad_l rounded up to even), conv row h =
    sum_e A_{h + e + e_min} @ W2[:, e, :].T  -- the index algebra of the kernel, without the GPU."""
    n, c, H, W = x.shape
    P = pad_l + (pad_l & 1)
    OH = (H + 2 * pad_t - kh) // 2 + 1
    OW = (W + 2 * pad_l - kw) // 2 + 1
    xp = np.zeros((n, c, H + 64, W + 64), np.float32)          # generous zero frame: index (row + 32, col + 32)
    xp[:, :, 32:32 + H, 32:32 + W] = x
    out = np.zeros((n, W2.shape[0], OH, OW), np.float32)
    for h in range(OH):
        for e in range(taps):
            t = h + e + e_min
            A = np.zeros((n, OW, 64), np.float32)
            for ph in range(2):
                for ci in range(c):
                    for j in range(8):
                        cols = 2 * np.arange(OW) + j - P + 32
                        A[:, :, (ph * 3 + ci) * 8 + j] = xp[:, ci, 2 * t + ph + 32, cols]
            out[:, :, h, :] += np.einsum('nwk,ok->now', A, W2[:, e, :].astype(np.float32))
    return out


@pytest.mark.parametrize('k,pad', [(7, 3), (5, 2), (3, 1), (6, 2), (4, 1)])
def test_stem_pool_weight_packing_reproduces_the_convolution(lib, k, pad):
    """The fused first-layer kernel's filter layout (ops.stem_pool_weight) and its packed-row algebra, emulated in
    numpy, must equal the oracle convolution (planer/util.py:17-44) -- this pins e_min / taps / column shift on the CPU."""
    from planer_b200 import ops
    rng = np.random.default_rng(k)
    x = rng.standard_normal((2, 3, 20, 24)).astype(np.float32)
    K = rng.standard_normal((64, 3, k, k)).astype(np.float32)
    e_min, taps, shift = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.plnr_stem_pool_geometry(k, pad, pad, ctypes.byref(e_min), ctypes.byref(taps), ctypes.byref(shift)) == 0
    assert shift.value == pad % 2 and 1 <= taps.value <= 4
    W2 = ops.stem_pool_weight(K, pad, pad).astype(np.float32)
    assert W2.shape == (64, taps.value, 64)
    got = _emulate_stem_contraction(x, W2, k, k, pad, pad, e_min.value, taps.value)
    ref = oracle.conv2d(x, K, None, 1, (2, 2), (1, 1), (pad,) * 4)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max()      # W2 is stored in fp16


def test_stem_pool_supported_truth_table(lib):
    names = ('dtype', 'c', 'h', 'w', 'cout', 'kh', 'kw', 'stride', 'pad_t', 'pad_l', 'pad_b', 'pad_r', 'act', 'pool_k',
             'pool_stride', 'pool_pad')
    base = dict(dtype=_capi.F16, c=3, h=224, w=224, cout=64, kh=7, kw=7, stride=2, pad_t=3, pad_l=3, pad_b=3, pad_r=3,
                act=_capi.ACT_RELU, pool_k=3, pool_stride=2, pool_pad=1)
    f = lambda **kw: lib.plnr_stem_pool_supported(*[{**base, **kw}[n] for n in names])
    assert f() == 1                                   # the ResNet stem
    assert f(h=64, w=64) == 1
    assert f(dtype=_capi.F32) == 0                    # fp32 runs the generic path
    assert f(act=_capi.ACT_NONE) == 0                 # only a ReLU makes the pooling pad neutral
    assert f(cout=32) == 0 and f(c=4) == 0 and f(stride=1) == 0
    assert f(pool_k=2) == 0 and f(pool_pad=0) == 0
    assert f(w=228) == 0                              # rows must be 16-byte multiples
    assert f(h=512, w=512) == 0                       # conv row wider than one 128-row MMA tile
    assert f(kh=9, kw=9, pad_t=4, pad_l=4, pad_b=4, pad_r=4) == 0


def test_resnet18_executor_patterns_are_planned():
    """The graph patterns the executor hands to single kernels exist in the compiled plan: input -> conv(relu) ->
    maxpool(3,2,1) with one consumer each, and gap -> flatten -> dense at the tail."""
    model, _ = zoo.resnet18(0)
    plan = P.compile_graph(model, {'x': (4, 3, 224, 224)})
    ops_ = [s.op for s in plan.steps]
    assert ops_[0] == 'conv' and ops_[1] == 'maxpool' and plan.steps[0].act == _capi.ACT_RELU
    assert plan.steps[1].attrs['w'] in ([3, 3], (3, 3)) and tuple(plan.steps[1].attrs['pads']) == (1, 1, 1, 1)
    assert ops_[-3:] == ['gap', 'flatten', 'dense'] or ops_[-4:-1] == ['gap', 'flatten', 'dense']


def test_unknown_operator_raises_by_name():
    model = {'input': ['x'], 'inits': [], 'layers': [['sm', 'logsoftmax', {}]], 'flow': [['x', ['sm'], 'y']]}
    with pytest.raises(NotImplementedError, match='logsoftmax'):
        P.compile_graph(model, {'x': (1, 10)})
    with pytest.raises(NotImplementedError, match='logsoftmax'):
        planer.Net().load_json(model['input'], model['inits'], model['layers'], model['flow'])
    # softmax is implemented along the stored innermost axis only: anything else is refused by name, not computed wrongly
    model['layers'][0][1] = 'softmax'
    model['layers'][0][2] = {'axis': 0}
    with pytest.raises(NotImplementedError, match='softmax'):
        P.compile_graph(model, {'x': (1, 10)})


def test_bad_pads_rejected_like_documented():
    model, _ = zoo.single_conv(8, 8, 3, pad=0)
    model['layers'][0][2]['pads'] = [0, 0, 1, 1]
    with pytest.raises(ValueError, match='undefined in the reference'):
        P.compile_graph(model, {'x': (1, 8, 8, 8)})


# ---------------------------------------------------------------------------------------------
# Net plumbing: our front-end driving the oracle's numpy operator table must be BIT-IDENTICAL to the
# reference Net (BASELINE config 1 "plumbing + bit-level correctness": IR parsing, blob slicing, flow
# chaining, liveness).  The table/array module are injected by the test; the package itself has no CPU path.
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize('name', ['readme_f32', 'readme_f16', 'resnet18_small_f32', 'yolov3_quarter_f32', 'decoder_f32',
                                  'decoder_f16', 'upsample_net_f32'])
def test_net_plumbing_bit_exact_with_injected_numpy_table(name):
    g = np.load(os.path.join(GOLD, 'graphs.npz'))
    model, blob, x, half = cases.make_graph_case(name)
    net = planer.Net(table=oracle.layer_map, array_module=np)
    net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
    net.load_weights(blob)
    if half:
        net.half()
    y = net(x.copy())
    ys = y if isinstance(y, tuple) else (y,)
    for i, t in enumerate(ys):
        assert cases.sample(t).tobytes() == g['%s.out%d' % (name, i)].tobytes()
    assert set(net.timer) >= {'conv'}
    # Net.map (a stream of batches) degrades to one call per batch off the GPU path: same results, same order
    xs = [x.copy(), x[:1].copy(), x.copy()]
    got = list(net.map(iter(xs)))
    for xi, gi in zip(xs, got):
        want = net(xi.copy())
        for a, b in zip(gi if isinstance(gi, tuple) else (gi,), want if isinstance(want, tuple) else (want,)):
            assert np.asarray(a).tobytes() == np.asarray(b).tobytes()


def test_c1_single_conv_plumbing_bit_exact():
    """BASELINE config 1 through the whole front-end (zoo IR -> Net -> op table) on numpy."""
    ops_gold = np.load(os.path.join(GOLD, 'ops.npz'))
    for name, pad in (('c1_conv_p0', 0), ('c1_conv_p1', 1)):
        kind, (x, K, B), kw = cases.make_case(name)
        model, _ = zoo.single_conv(3, 64, 3, 1, pad)
        blob = np.concatenate([K.reshape(-1).view(np.uint8), B.view(np.uint8)])
        net = planer.Net(table=oracle.layer_map, array_module=np)
        net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
        net.load_weights(blob)
        assert net(x.copy()).tobytes() == ops_gold[name].tobytes()


def test_model_files_roundtrip(tmp_path):
    """.json+.npy and .pla written by zoo.save_model carry the reference's format (planer/io.py:8-24,286)."""
    import zipfile
    model, blob = zoo.readme_net(0)
    zoo.save_model(str(tmp_path / 'm'), model, blob)
    zoo.save_model(str(tmp_path / 'p'), model, blob, pla=True)
    assert np.load(tmp_path / 'm.npy').dtype == np.uint8 and np.load(tmp_path / 'm.npy').ndim == 1
    body = json.load(open(tmp_path / 'm.json'))
    assert set(body) == {'input', 'inits', 'layers', 'flow'}
    with zipfile.ZipFile(tmp_path / 'p.pla') as z:
        assert sorted(z.namelist()) == ['p.json', 'p.npy']
    with contextlib.redirect_stdout(io.StringIO()) as out:
        assert planer.read_net(str(tmp_path / 'missing')) is None     # planer/io.py:30-31
    assert 'not found' in out.getvalue()
    if os.path.isdir('/root/reference/planer'):
        # the reference itself can read our files (only in the authoring container)
        import subprocess, sys
        code = ("import sys; sys.path.insert(0, '/root/reference'); import planer, numpy as np;"
                "net = planer.read_net(%r); print(net(np.zeros((1,3,8,8),'float32')).shape)" % str(tmp_path / 'p'))
        r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True,
                           env=dict(os.environ, HOME=str(tmp_path), PYTHONDONTWRITEBYTECODE='1'))
        assert '(1, 128, 8, 8)' in r.stdout, r.stderr[-500:]


# ---------------------------------------------------------------------------------------------
# multi-process path on CPU: gloo, world_size 2
# ---------------------------------------------------------------------------------------------

def _gloo_worker(rank, world, port, q):
    import torch.distributed as td
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    td.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model, blob = zoo.readme_net(0)
        total = blob.size
        got = dist.broadcast_host_blob(blob if rank == 0 else None, total)      # only rank 0 holds the bytes
        lo, hi = dist.shard_batch(5)
        # each rank runs ITS shard through the (injected numpy) front-end; no collective on the forward path
        net = planer.Net(table=oracle.layer_map, array_module=np)
        net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
        net.load_weights(got)
        x = np.random.default_rng(1).standard_normal((5, 3, 16, 16)).astype(np.float32)
        y = net(x[lo:hi].copy())
        t = dist.max_over_ranks(float(rank + 1))
        import zlib
        crcs = dist.gather_ints(zlib.crc32(got.tobytes()))                       # bench.py's load check: one CRC per rank
        q.put((rank, lo, hi, got.tobytes() == blob.tobytes(), y, t, crcs))
    finally:
        td.destroy_process_group()


def test_gloo_world2_blob_broadcast_and_batch_shard():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs: p.join(60)
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]
    assert all(r[3] for r in res), 'rank did not receive the exact blob bytes'
    assert all(r[5] == 2.0 for r in res)
    import zlib
    model, blob = zoo.readme_net(0)
    assert all(r[6] == [zlib.crc32(blob.tobytes())] * 2 for r in res), 'gather_ints: every rank must see both CRCs'
    x = np.random.default_rng(1).standard_normal((5, 3, 16, 16)).astype(np.float32)
    full = oracle.build_net(model, blob)(x.copy())
    assert np.array_equal(np.concatenate([res[0][4], res[1][4]]), full)      # shards concatenate to the full batch


def test_shard_batch_covers_everything():
    for n in (1, 7, 128, 1024):
        for w in (1, 2, 4, 8):
            parts = [dist.shard_batch(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def test_decoder_plan_expands_convtranspose_into_zero_stuff_plus_conv():
    """ConvTranspose2d (planer/layer.py:28-34) compiles to a zero-stuffing step + an ordinary stride-1
    conv step (flipped filter) whose epilogue still absorbs bias / relu / sigmoid; AveragePool is one HBM step."""
    model, _ = cases.get_model('decoder')
    gp = P.compile_graph(model, {'x': (2, 3, 32, 40)})
    assert gp.summary() == {'conv': 4, 'averagepool': 1, 'zero_stuff': 2, 'concat': 1}
    steps = {s.name: s for s in gp.steps}
    up, out = steps['up'], steps['out']
    assert up.attrs['flip'] and up.attrs['strides'] == (1, 1) and up.attrs['pads'] == (0, 0, 0, 0)
    assert up.fused == ['up', 'up.relu'] and up.act == 1 and up.bias is not None
    assert out.fused == ['out', 'out.sigmoid'] and out.act == 3
    st = steps['up.stuff']
    assert st.attrs == {'lo_h': 2, 'lo_w': 2, 'strides': (2, 2)}
    shp = {s.name: gp.values[s.out].shape for s in gp.steps}
    assert shp['up.stuff'] == (2, 64, 35, 43) and shp['up'] == (2, 32, 32, 40) and shp['out'] == (2, 3, 32, 40)
    # 2 * N * Cout * Cin * kh * kw * OH * OW of the stride-1 conv over the stuffed buffer (what the kernels execute)
    assert [n.flops for n in gp.nodes if n.name == 'up'] == [2 * 2 * 32 * 64 * 16 * 32 * 40]


# ---------------------------------------------------------------------------------------------
# sliding-window inference
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize('name', list(cases.TILE_CASES))
def test_tile_decorator_equals_the_reference(name):
    """planer_b200.util.tile vs the reference's decorator (planer/util.py:291-348) on the same seeded image and per-window
    function: bit-identical (same float32 arithmetic, same window grid, same uint16 blending weights) -- window by window
    and with all windows stacked on the batch axis (``batched=True``)."""
    import hashlib
    from planer_b200 import util
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'tile.npz'))
    img, kw, fn = cases.make_tile_case(name)
    assert hashlib.sha256(img.tobytes()).hexdigest() == str(gold[name + '.sha'])
    seen = []
    y = util.tile(progress=lambda i, n: seen.append((i, n)), **kw)(fn)(img.copy())
    ref = gold[name]
    assert y.shape == ref.shape and y.dtype == ref.dtype and y.tobytes() == ref.tobytes()
    if seen:
        assert seen == [(i + 1, seen[0][1]) for i in range(seen[0][1])]          # one progress call per window, in order
    calls = []

    def stacked(batch):
        calls.append(batch.shape[0])
        return np.stack([fn(b) for b in batch])
    yb = util.tile(batched=True, **kw)(stacked)(img.copy())
    assert len(calls) == 1 and calls[0] == max(len(seen), 1)
    assert yb.tobytes() == ref.tobytes()
    # per-call keyword overrides, like the reference's wrapper
    y2 = util.tile(window=9999)(fn)(img.copy(), **{**kw, 'progress': lambda *a: None})
    assert y2.tobytes() == ref.tobytes()


# ---------------------------------------------------------------------------------------------
# ONNX importer against files written by PyTorch's own exporter (oracle/gen_onnx_fixtures.py)
# ---------------------------------------------------------------------------------------------
ONNX_DIR = os.path.join(ROOT, 'tests', 'golden', 'onnx')


def _onnx_consts(model, blob):
    """Host copies of the small constant inits the planner needs as numbers (upsample / resize scales)."""
    out, s = {}, 0
    for name, shape, dt in model['inits']:
        nb = int(np.prod(shape)) * np.dtype(dt).itemsize
        if nb <= 64:
            out[name] = blob[s:s + nb].view(dt).reshape(shape)
        s += nb
    return out


@pytest.mark.parametrize('name', ['mini_resnet', 'mini_decoder'])
def test_onnx_import_reproduces_torch_through_the_oracle(name):
    """read_onnx (planer/io.py:53-287 restated, no onnx package) on a real exporter file: the IR, run by the numpy oracle,
    reproduces PyTorch's outputs; the IR follows the reference's conventions; the planner compiles it."""
    from planer_b200 import onnx_import
    model, blob = onnx_import.read_onnx(os.path.join(ONNX_DIR, name + '.onnx'))
    g = np.load(os.path.join(ONNX_DIR, name + '.npz'))
    y = oracle.build_net(model, blob)(g['x'].copy())
    ys = y if isinstance(y, tuple) else (y,)
    for i, t in enumerate(ys):
        assert t.shape == g['y%d' % i].shape and cases.rel_err(t, g['y%d' % i]) <= 1e-5
    # IR conventions of the reference importer
    assert model['input'][0] == 'x' and model['layers'][-1] == ['return', 'return', {}] and model['flow'][-1][2] == 'plrst'
    assert blob.dtype == np.uint8 and blob.size == sum(int(np.prod(s)) * np.dtype(d).itemsize for _, s, d in model['inits'])
    for xs, names, out in model['flow'][:-1]:
        assert len(names) == 1 and isinstance(out, str) and (isinstance(xs, str) or len(xs) > 1)      # io.py:70-73
    kinds = [l[1] for l in model['layers']]
    if name == 'mini_decoder':
        bn = [l[0] for l in model['layers'] if l[1] == 'batchnorm'][0]
        flow = [f for f in model['flow'] if f[1] == [bn]][0]
        assert flow[0][1].endswith('_invK') and flow[0][2].endswith('_invB')                           # io.py:85-90
        names = [i[0] for i in model['inits']]
        assert flow[0][1][:-5] in names                            # the original gamma stays in the blob (io.py:78,89)
        assert {'clip', 'resize', 'convtranspose', 'averagepool', 'hardsigmoid', 'softmax', 'leakyrelu'} <= set(kinds)
        clip = [l for l in model['layers'] if l[1] == 'clip'][0]
        assert clip[2] == {'min': 0.0, 'max': 6.0}                 # ReLU6: bounds arrive as inputs in opset 13
    gp = P.compile_graph(model, {'x': g['x'].shape}, _onnx_consts(model, blob))
    assert gp.summary().get('conv', 0) >= 4


def test_onnx_import_names_what_it_cannot_do(tmp_path):
    from planer_b200 import onnx_import
    # a one-node graph with an operator outside the table: field numbers per the ONNX schema (NodeProto.op_type = 4, ...)
    def msg(fields):
        out = b''
        for num, payload in fields:
            out += bytes([(num << 3) | 2, len(payload)]) + payload
        return out
    node = msg([(1, b'x'), (2, b'y'), (3, b'n0'), (4, b'Erf')])
    vi = lambda n: msg([(1, n)])
    graph = msg([(1, node), (11, vi(b'x')), (12, vi(b'y'))])
    with pytest.raises(NotImplementedError, match='Erf'):
        onnx_import.read_onnx(msg([(7, graph)]))
    with pytest.raises(ValueError, match='GraphProto'):
        onnx_import.read_onnx(b'')


# ---------------------------------------------------------------------------------------------
# the drop-in into the reference package (Level B): what install() rebinds, checked without a GPU
# ---------------------------------------------------------------------------------------------

def test_install_rebinds_the_reference_package():
    """planer_b200.install(planer) on the pip-installed, unmodified reference (baseline/_ref): the array module of
    util / layer / net / io and every hot-path entry of planer.layer.layer_map now point at the B200 implementation; the
    entries outside the hot path stay the reference's.  (The run on the GPU is tests/test_gpu_parity.py.)"""
    import sys
    ref_root = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'planer')):
        pytest.skip('baseline/_ref not present (created by __graft_entry__.build() where /root/reference exists)')
    os.environ.setdefault('HOME', '/tmp')
    sys.path.insert(0, ref_root)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import planer as ref
    finally:
        sys.path.remove(ref_root)
    saved = dict(ref.layer.layer_map)
    try:
        planer.install(ref)
        for mod in (ref.util, ref.layer, ref.net, ref.io):
            assert mod.np is planer.b200
        for kind, fn in planer.layer_map.items():
            assert ref.layer.layer_map[kind] is fn
        outside = set(saved) - set(planer.layer_map)
        assert outside and all(ref.layer.layer_map[k] is saved[k] for k in outside)
        assert ref.net.key is ref.layer.layer_map            # the reference's Net reads the patched table
    finally:
        ref.core(np, True)
        ref.layer.layer_map.clear()
        ref.layer.layer_map.update(saved)


def test_debug_trace_and_schedule_follow_the_reference_interpreter():
    """forward(debug=True) walks the schedule lowered at load_json: same trace lines, same liveness (a key is dropped
    after the last flow entry that reads it), same results as the fused-free reference loop (planer/net.py:37-72)."""
    model, blob, x, _ = cases.make_graph_case('readme_f32')
    net = planer.Net(table=oracle.layer_map, array_module=np)
    net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
    net.load_weights(blob)
    ops_ = net._schedule
    assert [o.layer for o in ops_] == ['conv', 'relu', 'pool', 'up', 'concat', 'sigmoid', 'return']
    assert ops_[0].ins == ('x', 'K', 'B') and not ops_[0].strict and ops_[1].ins == ('a',) and ops_[1].strict
    assert 'x' in ops_[0].dead and ops_[1].dead == ()          # chained layers never drop keys a second time
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        y = net.forward(x.copy(), debug=True)
    out = buf.getvalue()
    assert 'conv conv : ' in out and "\t-->  ['x', 'K', 'B'] :" in out and '\t<--  a :' in out
    g = np.load(os.path.join(GOLD, 'graphs.npz'))
    assert cases.sample(y[0]).tobytes() == g['readme_f32.out0'].tobytes()


# ---------------------------------------------------------------------------------------------
# shift-GEMM geometry (csrc/conv_shift.cu): the virtual grid, phase planes and tap tables the kernel uses, replayed in
# numpy against the oracle convolution -- pins the stride-2 plane algebra and the TMA start coordinates without a GPU
# ---------------------------------------------------------------------------------------------

def _shift_geometry(lib, xshape, cout, k, stride, pads, dil=1):
    n, c, h, w = xshape
    d = _capi.ConvDesc(_capi.F16, k[0], k[1], pads[0], pads[1], pads[2], pads[3], stride, stride, dil, dil, 1, 0)
    oh = (h + pads[0] + pads[2] - (k[0] - 1) * dil - 1 + stride) // stride
    ow = (w + pads[1] + pads[3] - (k[1] - 1) * dil - 1 + stride) // stride
    tx, ty = _capi.Tensor(256, n, h, w, c, c, 0), _capi.Tensor(256, n, oh, ow, cout, cout, 0)
    out = (ctypes.c_int * 128)()
    assert lib.plnr_debug_shift_geometry(ctypes.byref(d), ctypes.byref(tx), ctypes.byref(ty), out, 128) == 0
    return list(out), oh, ow


@pytest.mark.parametrize('cfg', [
    # n, c, h, w, cout, (kh, kw), stride, pads (t, l, b, r), dilation
    (2, 64, 12, 12, 8, (3, 3), 2, (1, 1, 1, 1), 1),       # ResNet down-sampling conv: four planes, 4 + 2 + 2 + 1 taps
    (1, 64, 13, 11, 8, (3, 3), 2, (1, 1, 1, 1), 1),       # odd sizes: the odd-phase planes are one row / column shorter
    (2, 64, 10, 14, 8, (1, 1), 2, (0, 0, 0, 0), 1),       # 1x1 / stride 2 shortcut conv: one plane, one tap
    (1, 64, 20, 22, 8, (3, 3), 2, (0, 0, 0, 0), 1),       # no padding
    (1, 64, 16, 16, 8, (5, 5), 2, (2, 2, 2, 2), 1),       # 5x5: 9 + 6 + 6 + 4 taps
    (1, 64, 12, 12, 8, (3, 3), 2, (1, 1, 0, 0), 1),       # top / left padding only
    (2, 64, 9, 10, 8, (3, 3), 1, (1, 1, 1, 1), 1),        # stride 1: one plane in filter order
    (1, 64, 12, 12, 8, (3, 3), 1, (2, 2, 2, 2), 2),       # stride 1, dilation 2
])
def test_shift_gemm_plane_geometry_reproduces_the_convolution(lib, cfg, monkeypatch):
    monkeypatch.setenv('PLNR_SHIFT_S2', '1')          # the stride-2 phase-plane path is opt-in (csrc/conv_shift.cu)
    n, c, h, w, cout, k, stride, pads, dil = cfg
    g, oh, ow = _shift_geometry(lib, (n, c, h, w), cout, k, stride, pads, dil)
    assert g[0] == 1, 'the shift kernel must apply to this problem'
    cs, Hv, Wv, halo, nplanes, ntaps = g[1:7]
    assert cs == stride and ntaps == k[0] * k[1]
    rng = np.random.default_rng(sum(cfg[:5]) + stride)
    x = rng.integers(-3, 4, (n, c, h, w)).astype(np.float64)           # small integers: exact arithmetic
    K = rng.integers(-2, 3, (cout, c, k[0], k[1])).astype(np.float64)
    ref = oracle.conv2d(x.astype(np.float32), K.astype(np.float32), None, 1, (stride, stride), (dil, dil), pads).astype(np.float64)
    M = n * Hv * Wv
    acc = np.zeros((M + halo + 1, cout))
    for pi in range(nplanes):
        w0, h0, first = g[7 + 3 * pi: 10 + 3 * pi]
        last = g[7 + 3 * (pi + 1) + 2] if pi + 1 < nplanes else ntaps
        # what the TMA boxes put into shared memory: virtual position (img, p', q') <- x[img, :, p' * cs + h0, q' * cs + w0], 0 outside
        plane = np.zeros((M + halo + 1, c))
        for img in range(n):
            for pv in range(Hv):
                ih = pv * cs + h0
                if not 0 <= ih < h:
                    continue
                for qv in range(Wv):
                    iw = qv * cs + w0
                    if 0 <= iw < w:
                        plane[(img * Hv + pv) * Wv + qv] = x[img, :, ih, iw]
        for t in range(first, last):
            aoff, wt = g[19 + 2 * t], g[20 + 2 * t]
            assert 0 <= aoff <= halo
            r, s = divmod(wt, k[1])
            shifted = np.zeros_like(plane)
            shifted[:M + halo + 1 - aoff] = plane[aoff:]
            acc += shifted @ K[:, :, r, s].T                           # one tcgen05.mma chain: D += A(shifted view) * W_tap^T
    got = acc[:M].reshape(n, Hv, Wv, cout)[:, :oh, :ow].transpose(0, 3, 1, 2)
    assert np.array_equal(got, ref)


def test_shift_gemm_rejects_what_it_cannot_do(lib, monkeypatch):
    assert _shift_geometry(lib, (1, 64, 12, 12), 8, (3, 3), 2, (1, 1, 1, 1), 1)[0][0] == 0      # stride 2 is opt-in
    monkeypatch.setenv('PLNR_SHIFT_S2', '1')
    assert _shift_geometry(lib, (1, 64, 12, 12), 8, (3, 3), 2, (1, 1, 1, 1), 1)[0][0] == 1
    for cfg in [((1, 64, 12, 12), 8, (3, 3), 3, (1, 1, 1, 1), 1),     # stride 3
                ((1, 64, 12, 12), 8, (3, 3), 2, (2, 2, 2, 2), 2),     # stride 2 with dilation
                ((1, 48, 12, 12), 8, (3, 3), 1, (1, 1, 1, 1), 1),     # Cin % 64 != 0
                ((1, 64, 12, 300), 8, (3, 3), 2, (1, 1, 1, 1), 1)]:   # 2 * Wv > 256: the TMA box limit
        assert _shift_geometry(lib, *cfg)[0][0] == 0


def test_zero_copy_concat_placement_and_liveness():
    """P.place_concat_inputs: YOLOv3's four concat inputs (two nearest upsamples, two route convolutions that other
    convolutions read too) are produced inside the concat buffers; the buffer of a concat is then live from its first
    producer to the last reader of the concat or of any slice, and nothing else may take it in between."""
    model, _ = cases.get_model('yolov3_quarter')
    consts = {n: np.array([1, 1, 2, 2], np.float32) for n, s, d in model['inits'] if 'scales' in n}
    gp = P.compile_graph(model, {'x': (1, 3, 64, 64)}, consts)
    n_concat_inputs = sum(len(s.ins) for s in gp.steps if s.op == 'concat')
    assert P.place_concat_inputs(gp) == n_concat_inputs == 4
    P.assign_buffers(gp, 2, lambda v: gp.values[v].shape[1])
    sroot = lambda v: P._storage_root(gp.values, v)
    for s in gp.steps:
        if s.op == 'concat':
            off = 0
            for i in s.ins:
                v = gp.values[P._root(gp.values, i)]
                assert v.slice_of == (P._root(gp.values, s.out), off) and sroot(i) == sroot(s.out)
                assert P._root(gp.values, i) not in gp.buffer_of       # no buffer of its own
                off += v.shape[1]
    last = {}
    for pos, s in enumerate(gp.steps):
        for r in s.reads():
            last[sroot(r)] = pos
    tenant = {}
    for pos, s in enumerate(gp.steps):
        o = sroot(s.out)
        if o in gp.buffer_of:
            b = gp.buffer_of[o]
            prev = tenant.get(b)
            if prev is not None and prev != o:
                assert last.get(prev, -1) < pos, (s.name, prev, o)
            tenant[b] = o
    # a value with a reader that cannot take a strided view (elementwise add) is not placed
    b = zoo._Builder(3)
    a1 = b.conv('x', 16, 16, 3, 1, 1, name='a1')
    a2 = b.conv('x', 16, 16, 3, 1, 1, name='a2')
    cat = b.op('concat', {'axis': 1}, [a1, a2], name='cat')
    s_ = b.op('add', {}, [a1, 'x'], name='add')
    model2, _ = b.finish(['x'], [cat, s_])
    gp2 = P.compile_graph(model2, {'x': (1, 16, 8, 8)})
    assert P.place_concat_inputs(gp2) == 1
    assert gp2.values[P._root(gp2.values, [s for s in gp2.steps if s.name == 'a2'][0].out)].slice_of is not None


def test_fp16_split_arithmetic_reproduces_a_float32_dot_product():
    """The algebra of csrc/split_f32.cu, emulated in numpy (no GPU): both operands scaled by a power of two into
    [2^13, 2^14), split into fp16 (hi, lo), and  xh.wh + xh.wl + xl.wh  accumulated in fp32 reproduce a float32 dot product to
    ~1e-6 of its range for tensors of any magnitude (1e-5 .. 3e3 here) -- the bar of the float32 path is 1e-3."""
    rng = np.random.default_rng(7)

    def prescale(a):
        m = float(np.abs(a).max())
        return 2.0 ** (13 - int(np.floor(np.log2(m)))) if m > 0 else 1.0

    def split(a, s):
        v = (a * np.float32(s)).astype(np.float32)
        hi = v.astype(np.float16)
        lo = (v - hi.astype(np.float32)).astype(np.float16)
        assert np.isfinite(hi.astype(np.float32)).all()
        return hi.astype(np.float32), lo.astype(np.float32)

    for mag in (1.0, 1e-5, 3e3):
        x = (rng.standard_normal((64, 576)) * mag).astype(np.float32)
        w = (rng.standard_normal((576, 32)) * np.sqrt(2.0 / 576)).astype(np.float32)
        sx, sw = prescale(x), prescale(w)
        xh, xl = split(x, sx)
        wh, wl = split(w, sw)
        acc = (xh @ wh + xh @ wl + xl @ wh).astype(np.float32)              # every fp16 x fp16 product is exact in fp32
        got = acc * np.float32(1.0 / (sx * sw))
        ref = x.astype(np.float64) @ w.astype(np.float64)
        assert np.abs(got - ref).max() / np.abs(ref).max() <= 2e-6, mag
        # dropping the low halves (plain fp16 operands) is three orders of magnitude worse
        assert np.abs(xh @ wh / (sx * sw) - ref).max() / np.abs(ref).max() >= 1e-4


def test_tile_decorator_random_cases_equal_the_reference_itself():
    """planer_b200.util.tile against the UNMODIFIED reference's decorator (baseline/_ref, planer/util.py:291-348) on random image
    sizes, windows, margins, sampling factors and per-window functions: bit-identical, also with ``batched=True``."""
    ref_root = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'planer')):
        pytest.skip('baseline/_ref is not in this checkout')
    import sys
    if not os.access(os.path.expanduser('~'), os.W_OK):
        os.environ['HOME'] = '/tmp'
    sys.path.insert(0, ref_root)
    try:
        import planer as ref
    finally:
        sys.path.remove(ref_root)
    from planer_b200 import util
    rng = np.random.default_rng(11)
    quiet = lambda *a: None
    skipped = 0
    for case in range(60):
        rgb = bool(rng.integers(0, 2))
        shape = (int(rng.integers(20, 200)), int(rng.integers(20, 200))) + ((3,) if rgb else ())
        img = (rng.standard_normal(shape) * 10).astype(np.float32)
        kind = str(rng.choice(['same', 'up2', 'down2'] + (['gray'] if rgb else [])))
        kw = dict(window=int(rng.choice([16, 24, 32, 48, 64, 256])), margin=float(rng.choice([0.1, 0.25, 0.3])) if rng.integers(0, 2) else int(rng.integers(1, 8)),
                  glob=int(rng.choice([1, 8, 16])), progress=quiet)
        if rng.integers(0, 3) == 0:
            kw['sample'] = float(rng.choice([0.5, 0.75, 1.5]))
        fn = cases.tile_fn(kind)
        try:
            want = ref.util.tile(**kw)(fn)(img.copy())
        except Exception:          # parameter combinations the reference itself cannot run (odd extents under a /2 function, ...)
            skipped += 1
            continue
        got = util.tile(**kw)(fn)(img.copy())
        assert got.shape == want.shape and got.dtype == want.dtype, (case, kw, shape, kind)
        assert np.array_equal(got, want), (case, kw, shape, kind)
        stacked = lambda ws, fn=fn: np.stack([fn(w_) for w_ in ws])
        assert np.array_equal(util.tile(batched=True, **kw)(stacked)(img.copy()), want), (case, kw, shape, kind)
    assert skipped <= 30, skipped


def test_pool_fold_applies_only_to_whole_32_position_parts(lib, monkeypatch):
    """plnr_conv2d_pool_parts (host logic of plnr_epilogue.pool_sum): the GlobalAveragePool fold needs the fp16 stride-1
    shift-GEMM kernel, output channels in whole 32-column chunks and a padded image grid of whole 32-position parts, so that
    every TMEM lane quarter of a tile lies inside one image."""
    def parts(xs, cout, k, stride=1, pad=None, dtype=np.float16, groups=1):
        n, c, h, w = xs
        pad = k // 2 if pad is None else pad
        oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        d = _capi.ConvDesc(_capi.dtype_code(np.dtype(dtype)), k, k, pad, pad, pad, pad, stride, stride, 1, 1, groups, 0)
        tx = _capi.Tensor(256, n, h, w, c, c, 0)
        ty = _capi.Tensor(256, n, oh, ow, cout, cout, 0)
        return lib.plnr_conv2d_pool_parts(ctypes.byref(d), ctypes.byref(tx), ctypes.byref(ty))
    assert parts((128, 512, 7, 7), 512, 3) == 2               # ResNet-18 at 224 x 224: (7 + 1) x (7 + 1) = 64 positions
    assert parts((4, 64, 15, 15), 32, 3) == 8
    assert parts((2, 128, 31, 31), 64, 3) == 32
    assert parts((6, 64, 8, 8), 256, 1) == 2                  # 1 x 1 without padding: the grid is the image
    assert parts((2, 512, 8, 8), 512, 3) == 0                 # 256 x 256 inputs: 81 positions per image
    assert parts((2, 512, 7, 7), 1000, 3) == 0                # output channels not in whole 32-column chunks
    assert parts((2, 48, 7, 7), 64, 3) == 0                   # input channels not in 64-channel chunks: another kernel
    assert parts((2, 64, 14, 14), 64, 3, stride=2) == 0       # strided: the im2col kernel
    assert parts((2, 64, 7, 7), 64, 3, dtype=np.float32) == 0
    assert parts((2, 64, 7, 7), 64, 3, groups=2) == 0
    monkeypatch.setenv('PLNR_NO_POOL_FOLD', '1')
    assert parts((128, 512, 7, 7), 512, 3) == 0
