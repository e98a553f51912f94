"""Headline benchmark: ResNet-18 224x224 forward images/sec on N B200s (BASELINE.json metric), fp16,
batch 128 per GPU (configs[2]; configs[4] = 8 x 128 is the same workload weak-scaled to 8 GPUs).
``--config yolov3`` runs BASELINE configs[3] (YOLOv3-416 fp16, batch 32) through the same code and prints the same line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config resnet18|yolov3]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  A "step" is one forward of the whole graph over one batch of synthetic NCHW input.

  * Before anything is timed the net is CHECKED in this process: its output on the committed golden input must match the
    unmodified reference's output (tests/golden/graphs.npz) within the north star's tolerance; with N > 1 every rank
    reports the CRC-32 of the weight blob it received over NCCL and of its logits on that shared input, and rank 0
    requires them all equal.  A failed check aborts the bench.
  * `value`: device-timed (CUDA events on the library stream), inputs resident in HBM, rotating over buffers > L2.
  * `e2e`: the same metric through the public API from HOST memory -- ``for y in net.map(batches)`` with one pinned host
    batch uploaded and the result read back for EVERY step inside the timed region, at least 100 batches and 0.5 s
    whatever --steps says.  The headline `e2e.value` feeds uint8 images (what an image pipeline holds; the first layer
    converts on the fly -- the numpy reference computes on ``x.astype(float16)`` for such an input); `e2e.fp16_host` is
    the same with fp16 host batches (twice the PCIe bytes) and `e2e.blocking_call` the reference-shaped ``y = net(x)``.
  * `roofline`: whole-step tensor roofline (every launch of the step is a conv / dense kernel): algorithmic FLOPs of the
    plan / the graph-timed step, against the sustained bf16 peak of MEASURED_PEAKS.json; `roofline.families` splits the
    step by kernel family (live un-graphed CUDA-event pairs, scaled to the graph-timed step).
  * `cpu_baseline` and ``--impl reference``: the UNMODIFIED reference (baseline/_ref, installed from /root/reference by
    pip) on the host cores; the numpy restatement in oracle/ is the fallback when that directory is missing.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = 'images/s'
N_INPUT_BUFFERS = 8
CONFIGS = {
    # 8 x 38.5 MB = 308 MB of rotating inputs > 126 MB L2
    'resnet18': dict(metric='resnet18_224_fwd_images_per_sec', hw=224, batch=128, golden='resnet18_f32_n1', tol=1e-2,
                     workload='ResNet-18 224x224 fp16 forward, batch %d per GPU (BASELINE configs[2]; x8 = configs[4])',
                     cpu_batch=8),
    # 8 x 33 MB = 266 MB of rotating inputs
    'yolov3': dict(metric='yolov3_416_fwd_images_per_sec', hw=416, batch=32, golden='yolov3_416_f32_n1', tol=1e-2,
                   workload='YOLOv3-416 (Darknet-53 + 3 heads) fp16 forward, batch %d per GPU (BASELINE configs[3])',
                   cpu_batch=1),
}


def build_model(config):
    from planer_b200 import zoo
    return zoo.resnet18(0) if config == 'resnet18' else zoo.yolov3(0)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'tflops_burst': p['bf16_tflops'], 'tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'hbm_gbs': p['hbm_gbs'], 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'hbm_gbs': 6650.0,
            'source': 'fallback (B200_PROFILING.md)'}


def lib_sha256():
    from planer_b200 import _capi
    h = hashlib.sha256()
    with open(_capi.LIB_PATH, 'rb') as f:
        for chunk in iter(lambda: f.read(1 << 20), b''):
            h.update(chunk)
    return h.hexdigest()


def src_sha256():
    """Hash of the CUDA sources + the C header the library is built from: identifies the build when the .so itself was rebuilt
    on another box (-lineinfo embeds the source paths, so the binary hash then differs for identical code)."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, 'planer_b200', 'csrc')
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(('.cu', '.cuh')))
    for fn in files + [os.path.join(ROOT, 'include', 'planer_b200.h')]:
        h.update(os.path.basename(fn).encode())
        with open(fn, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_median': float(np.median(pw)) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on the host cores
# ------------------------------------------------------------------------------------------------------------------
def reference_net(model, blob):
    """(net, kind): the UNMODIFIED reference's Net from baseline/_ref ('reference'), else the numpy restatement ('port')."""
    ref_root = os.path.join(ROOT, 'baseline', '_ref')
    if os.path.isdir(os.path.join(ref_root, 'planer')):
        if not os.access(os.path.expanduser('~'), os.W_OK):
            os.environ['HOME'] = '/tmp'                 # the reference creates ~/.planer_zoo at import (planer/__init__.py:50-51)
        sys.path.insert(0, ref_root)
        try:
            import planer
            planer.core(np, True)
            net = planer.Net()
            net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
            net.load_weights(blob)
            return net, 'reference'
        except Exception as e:                          # pragma: no cover
            print('[bench] reference import failed (%r): using the oracle port' % (e,), file=sys.stderr)
        finally:
            sys.path.remove(ref_root)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import planer_oracle as oracle
    return oracle.build_net(model, blob), 'port'


def cpu_baseline_run(config, steps, warmup, max_sec=25.0):
    """Reference numpy path on the host cores, fp32 (numpy fp16 matmul has no BLAS: 18 s/img), bounded sample."""
    cfg = CONFIGS[config]
    model, blob = build_model(config)
    net, kind = reference_net(model, blob)
    batch = cfg['cpu_batch']
    x = np.random.default_rng(1).standard_normal((batch, 3, cfg['hw'], cfg['hw'])).astype(np.float32)
    cores = len(os.sched_getaffinity(0))
    # torchrun exports OMP_NUM_THREADS=1, which OpenBLAS obeys at import: give the reference path every host core back
    import contextlib
    try:
        from threadpoolctl import threadpool_limits
        blas = threadpool_limits(limits=cores)
    except ImportError:
        blas = contextlib.nullcontext()
    with blas:
        for _ in range(warmup):
            net(x.copy())
        t0, done = time.perf_counter(), 0
        for _ in range(steps):
            net(x.copy())
            done += 1
            if time.perf_counter() - t0 > max_sec:
                break
        dt = time.perf_counter() - t0
    return {'value': batch * done / dt, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': '%s fp32 (numpy fp16 matmul has no BLAS) batch %d x %d forwards, numpy %s, BLAS threads = all %d cores%s'
                      % (config, batch, done, np.__version__, cores,
                         ', unmodified reference from baseline/_ref' if kind == 'reference' else ', oracle/planer_oracle.py')}, dt / done, done


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    steps, warm = max(1, min(args.steps, 20)), max(1, min(args.warmup, 2))
    cb, sec, done = cpu_baseline_run(args.config, steps, warm)
    line = {'impl': 'reference', 'metric': cfg['metric'], 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': done, 'warmup': warm, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': '%s forward, reference numpy path on host cores, bounded sample: batch %d per step'
                                   % (args.config, cfg['cpu_batch'])},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    _emit(line)


# ------------------------------------------------------------------------------------------------------------------
# device timing helpers
# ------------------------------------------------------------------------------------------------------------------
def per_kernel_profile(net, x_dev, reps=3):
    """Un-graphed pass with a CUDA event pair around every launch: best-of-`reps` duration per launch (ms)."""
    import ctypes as C
    from planer_b200 import backend as B
    ex = net.executor([x_dev.shape], [x_dev.dtype])
    lib, ctx = B.lib(), B.ctx()
    fns = [lambda: ex._load_inputs([x_dev])] + list(ex.launches)
    n = len(fns)
    best, kernels = [float('inf')] * n, [''] * n
    for _ in range(reps):
        evs = []
        for i, fn in enumerate(fns):
            a, b = C.c_void_p(), C.c_void_p()
            lib.plnr_event_create(C.byref(a)); lib.plnr_event_create(C.byref(b))
            lib.plnr_event_record(ctx, a)
            fn()
            kernels[i] = B.last_kernel()
            lib.plnr_event_record(ctx, b)
            evs.append((a, b))
        B.synchronize()
        for i, (a, b) in enumerate(evs):
            ms = C.c_float()
            lib.plnr_event_elapsed_ms(a, b, C.byref(ms))
            best[i] = min(best[i], ms.value)
            lib.plnr_event_destroy(a); lib.plnr_event_destroy(b)
    fused = [f for f in ex.fused_stems.values()]
    in_kind = 'conv' if fused else 'input'
    in_name = ('+'.join(fused[0]['conv'].fused + (fused[0]['pool'].fused if fused[0]['pool'] else [])) +
               ' (one kernel, at input time)') if fused else 'input layout'
    return [{'kind': k, 'name': nm, 'ms': t, 'kernel': kn}
            for k, nm, t, kn in zip([in_kind] + list(ex.kinds), [in_name] + list(ex.names), best, kernels)]


def time_graph_steps(net, xs, steps):
    """ms per forward over `steps` back-to-back forwards (device events on the library stream)."""
    import torch
    from planer_b200 import backend as B
    stream = B.stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    B.synchronize()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(steps):
            net.forward(xs[i % len(xs)])
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def parity_check(net, config, half=True):
    """The timed net on the committed golden input vs the unmodified reference's output (tests/golden/graphs.npz)."""
    from tests import cases
    cfg = dict(CONFIGS[config])
    if not half:
        cfg['tol'] = 1e-3                              # north star: 1e-3 for fp32, 1e-2 for fp16
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'graphs.npz'))
    name = cfg['golden']
    _, shape, _ = cases.GRAPH_CASES[name]
    x = np.random.default_rng(1).standard_normal(shape).astype(np.float16 if half else np.float32)   # cases.make_graph_case, in the net's dtype
    y = net(x)
    ys = y if isinstance(y, tuple) else (y,)
    assert len(ys) == int(gold[name + '.nout']), 'output count differs from the reference'
    worst = 0.0
    for i, t in enumerate(ys):
        assert tuple(t.shape) == tuple(gold['%s.shape%d' % (name, i)]), 'output shape differs from the reference'
        ref, scale = gold['%s.out%d' % (name, i)], float(gold['%s.absmax%d' % (name, i)])
        worst = max(worst, float(np.abs(cases.sample(t).astype(np.float64) - ref.astype(np.float64)).max() / scale))
    if not worst <= cfg['tol']:
        raise SystemExit('[bench] PARITY FAILURE: %s %s vs reference golden %s: range-relative error %.3e > %.0e'
                         % (config, 'fp16' if half else 'fp32', name, worst, cfg['tol']))
    import zlib
    crc = zlib.crc32(b''.join(np.ascontiguousarray(t).tobytes() for t in ys))
    return {'golden': 'tests/golden/graphs.npz:' + name, 'rel_err': worst, 'tol': cfg['tol']}, crc


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything a library prints on fd 1 while the bench runs (NCCL's version banner, for one) goes to stderr: stdout carries
    exactly ONE line, the JSON result."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def _bind_numa(local):
    """Pin this rank's host threads to the CPUs of its GPU's NUMA node (pinned buffers are then allocated there by first
    touch): with 8 ranks feeding 8 GPUs the host -> device copies otherwise cross sockets.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = '/sys/bus/pci/devices/%s/numa_node' % bus.lower()[-12:]
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {'numa_node': node, 'cpus': len(allowed)}
    except Exception:
        return None
    return None


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='resnet18', choices=sorted(CONFIGS))
    ap.add_argument('--batch', type=int, default=None, help='images per GPU (default: the BASELINE config: 128 / 32)')
    ap.add_argument('--dtype', default='f16', choices=['f16', 'f32'],
                    help='f32 = BASELINE configs[1] arithmetic (CUDA-core FFMA path; roofline against the FFMA peak)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer legs (profiling runs)')
    ap.add_argument('--min-warm-sec', type=float, default=1.5, help='keep warming up at least this long (clock sampling)')
    ap.add_argument('--dump', default=None, help='write the per-kernel table to this JSON file')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank)
    cfg = CONFIGS[args.config]
    batch = args.batch or cfg['batch']

    numa = _bind_numa(local) if world > 1 else None
    import torch
    import planer_b200 as planer
    from planer_b200 import dist, backend as B
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    warmup, steps = max(args.warmup, 3), args.steps
    planer.core(planer.b200)
    B.init(local)

    # ---- model: rank 0 builds the blob, ONE NCCL broadcast hands it to the other ranks ----
    model, blob = build_model(args.config)
    net = planer.Net()
    net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
    net.load_weights(blob if rank == 0 else None) if world > 1 else net.load_weights(blob)
    import zlib
    n_init = sum(int(np.prod(s)) * np.dtype(d).itemsize for _, s, d in model['inits'])
    file_crc = zlib.crc32(np.ascontiguousarray(blob).reshape(-1).view(np.uint8)[:n_init].tobytes())
    blob_crcs = dist.gather_ints(net.blob_crc32())
    half = args.dtype == 'f16'
    if half:
        net.half()
    del blob
    npdt = np.float16 if half else np.float32

    # ---- checks before timing: parity against the reference's golden output; all ranks hold the same weights and
    #      produce the same logits on the shared golden input ----
    parity, logit_crc = parity_check(net, args.config, half)
    logit_crcs = dist.gather_ints(logit_crc)
    if rank == 0:
        if any(c != file_crc for c in blob_crcs):
            raise SystemExit('[bench] weight broadcast check FAILED: CRC-32 per rank %s, file %d' % (blob_crcs, file_crc))
        if any(c != logit_crcs[0] for c in logit_crcs):
            raise SystemExit('[bench] rank outputs differ on the shared input: CRC-32 per rank %s' % logit_crcs)
    checks = {'parity': parity, 'ranks_checked': len(blob_crcs), 'weight_blob_crc32_equal_on_all_ranks': True,
              'logits_crc32_equal_on_all_ranks': True}

    rng = np.random.default_rng(100 + rank)
    shape = (batch, 3, cfg['hw'], cfg['hw'])
    hosts = [rng.standard_normal(shape).astype(npdt) for _ in range(2)]
    xs = [B.asarray(hosts[i % 2] if i < 2 else np.roll(hosts[i % 2], i, axis=0)) for i in range(N_INPUT_BUFFERS)]
    B.synchronize()

    ex = net.executor([shape], [npdt])
    flops = ex.plan.flops
    # clocks are sampled from the start of the warm-up to the end of the timed region: the same load throughout.
    # The warm-up runs at least W steps AND at least ~1.5 s so that nvidia-smi (200 ms period) sees the loaded state.
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t_w, i = time.perf_counter(), 0
    while i < warmup or time.perf_counter() - t_w < args.min_warm_sec:
        net.forward(xs[i % N_INPUT_BUFFERS])
        i += 1
        if i % 50 == 0:
            B.synchronize()
    B.synchronize()
    warm_run = i

    # ---- timed region: device events on the library stream, barrier + sync on both sides ----
    stream = B.stream()
    dist.barrier(); torch.cuda.synchronize()
    l0 = B.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(steps):
            net.forward(xs[i % N_INPUT_BUFFERS])
        e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    ms_total = dist.max_over_ranks(e0.elapsed_time(e1))
    launches = B.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    value = world * batch * steps / (ms_total / 1e3)

    # ---- e2e: public API with host buffers (pinned), H2D + forward + D2H of EVERY step inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(net, hosts, batch, world, ms_total / steps, torch, dist, B)

    if rank != 0:
        return
    # ---- roofline: whole step = every conv / dense launch (tcgen05 kernels); per-family split from live event pairs ----
    table = per_kernel_profile(net, xs[0])
    all_ms = sum(r['ms'] for r in table)
    ms_step = ms_total / steps
    pk = peaks()
    split = (not half) and ex.split_convs > 0
    if split:
        # fp32 on the tensor pipe (csrc/split_f32.cu): every fp32 product is THREE fp16 tensor-core products, so the ceiling
        # for ALGORITHMIC fp32 FLOPs is a third of the fp16 tensor peak
        pk = dict(pk, tflops_sustained=pk['tflops_sustained'] / 3, tflops_burst=pk['tflops_burst'] / 3,
                  source=pk['source'] + ' / 3 (fp32 as three fp16 products per multiply)')
    elif not half:
        # fp32 on the CUDA cores (PLNR_F32_TENSOR=0): 148 SMs x 128 FFMA/clk x 2 FLOP at the clock sampled under this load
        mhz = (clk or {}).get('sm_mhz') or 1965.0
        ffma = B.device_info()['sm_count'] * 128 * 2 * mhz * 1e6 / 1e12
        pk = dict(pk, tflops_sustained=ffma, tflops_burst=ffma, source='FFMA peak = SMs x 128 x 2 x sampled SM clock (%.0f MHz)' % mhz)
    achieved = flops / (ms_step / 1e3) / 1e12
    fl_by_name = {n.name: n.flops for n in ex.plan.nodes}
    fam = {}
    for r in table:
        parts = [p.strip('()') for p in r['name'].replace(' (one kernel, at input time)', '').split('+')]
        f = sum(fl_by_name.get(p, 0) for p in parts)
        r['gflop'] = f / 1e9
        key = family_of(r, ex)
        d = fam.setdefault(key, {'family': key, 'launches': 0, 'ms_ungraphed': 0.0, 'gflop': 0.0})
        d['launches'] += 1; d['ms_ungraphed'] += r['ms']; d['gflop'] += f / 1e9
    families = []
    for d in fam.values():
        ms = d['ms_ungraphed'] * ms_step / all_ms          # scaled so that the families sum to the graph-timed step
        families.append({'family': d['family'], 'launches': d['launches'], 'ms': ms, 'share_of_step': ms / ms_step,
                         'gflop': d['gflop'], 'tflops': d['gflop'] / ms if ms > 0 else None,
                         'frac_of_sustained_peak': (d['gflop'] / ms) / pk['tflops_sustained'] if ms > 0 else None})
    families.sort(key=lambda d: -d['ms'])
    traffic, traffic_src = None, None
    sha = lib_sha256()
    cands = sorted(f for f in os.listdir(os.path.join(ROOT, 'profiles')) if f.endswith('_step_traffic.json'))
    if cands and batch == cfg['batch']:
        with open(os.path.join(ROOT, 'profiles', cands[-1])) as f:
            tj = json.load(f)
        if (tj.get('lib_sha256') == sha or tj.get('src_sha256') == src_sha256()) and tj.get('config', 'resnet18') == args.config:
            traffic = sum(int((k['dram_read_mb'] + k['dram_write_mb']) * 1e6) for k in tj['kernels'])
            traffic_src = ('dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one step, ncu --set full of '
                           'THIS build (profiles/%s: same library or same CUDA sources, lib sha256 %s)' % (cands[-1], sha[:12]))
        else:
            traffic_src = ('null: the newest committed capture (profiles/%s) is of another build of libplaner_b200.so'
                           % cands[-1])
    roofline = {'bound': 'tensor' if (half or split) else 'ffma',
                'kernel': 'whole step: every launch is a %s conv / dense kernel (%d conv+dense layers in %d launches)'
                          % ('tcgen05' if half else ('tcgen05 split-fp16 (fp32 operands as hi+lo fp16 pairs, fp32 accumulate; + '
                                                     'its absmax / split launches)' if split else 'CUDA-core FFMA'),
                             sum(1 for n in ex.plan.nodes if n.kind in ('conv', 'dense')), len(table)),
                'achieved': achieved, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['tflops_sustained'], 'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': pk['source'] + ', sustained figure (kernels timed inside a long step)',
                'frac_of_burst_peak': achieved / pk['tflops_burst'],
                'flops_per_step': flops, 'ms_per_step': ms_step,
                'algorithmic_bytes_per_step': ex.algorithmic_bytes(),
                'families': families, 'lib_sha256': sha}
    if args.dump:
        for r in table:
            r['tflops'] = r['gflop'] / r['ms'] if r['ms'] > 0 and r.get('gflop') else None
        with open(args.dump, 'w') as f:
            json.dump({'batch': batch, 'config': args.config, 'ms_per_step_graph': ms_step, 'lib_sha256': sha, 'table': table}, f, indent=1)

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb, _, _ = cpu_baseline_run(args.config, steps=8, warmup=1, max_sec=20.0)

    line = {'metric': cfg['metric'], 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': (cfg['workload'] % batch).replace('fp16', 'fp16' if half else 'fp32'), 'global_batch': world * batch,
                       'parallelism': 'dp%d batch split, no forward collective' % world,
                       'l2': 'inputs rotate over %d device buffers (%.0f MB > 126 MB L2)' % (N_INPUT_BUFFERS, N_INPUT_BUFFERS * hosts[0].nbytes / 1e6),
                       'launch': '%s + 1 CUDA graph (%d fused kernels) per step' % (
                           '1 input-time kernel (fused first layer)' if ex.fused_stems else 'input layout kernel', len(ex.launches)),
                       'warmup_steps_run': warm_run,
                       'warmup_note': 'at least --warmup steps AND %.1f s so that nvidia-smi (200 ms period) samples the loaded state' % args.min_warm_sec,
                       'numa_binding': numa},
            'clocks': clk, 'gpu_launches': int(launches), 'checks': checks,
            'e2e': e2e, 'roofline': roofline, 'cpu_baseline': cb}
    _emit(line)


def family_of(row, ex):
    """Kernel family of one launch of the step: the name the library reports for the kernel it picked."""
    return row.get('kernel') or row['kind']


def run_e2e(net, hosts, batch, world, ms_step, torch, dist, B):
    """Host-buffer legs.  Every leg: warm-up, barrier, wall clock around >= 100 batches and >= 0.5 s, max over ranks."""
    n_batches = max(100, int(0.5 / max(ms_step / 1e3, 1e-6)) + 1)
    n_batches = min(n_batches, 4000)
    rng = np.random.default_rng(7)

    def pin(a):
        p = B.pinned_empty(a.shape, a.dtype)
        p[...] = a
        return p

    def leg_map(bufs, n, copy=True):
        def feed(k):
            for i in range(k):
                yield bufs[i % len(bufs)]
        for y in net.map(feed(6), copy=copy):
            pass
        dist.barrier(); torch.cuda.synchronize()
        t0, nout, last = time.perf_counter(), 0, None
        for y in net.map(feed(n), copy=copy):
            last = y
            nout += (y[0] if isinstance(y, tuple) else y).shape[0]
        torch.cuda.synchronize()
        sec = dist.max_over_ranks(time.perf_counter() - t0)
        assert nout == batch * n
        return world * batch * n / sec, last

    u8 = [pin(rng.integers(0, 256, hosts[0].shape, dtype=np.uint8)) for _ in range(2)]
    v_u8, y = leg_map(u8, n_batches)
    d2h = int(sum(t.nbytes for t in (y if isinstance(y, tuple) else (y,))))
    f16 = [pin(h) for h in hosts]
    v_f16, _ = leg_map(f16, n_batches)
    v_views = None
    if d2h > (4 << 20):             # large results: the host memcpy into fresh arrays dominates; report the view-yielding call too
        v_views, _ = leg_map(u8, n_batches, copy=False)
    # the blocking reference-shaped call: upload + forward + download per call, no overlap between calls
    nb = max(20, n_batches // 4)
    for i in range(3):
        net(u8[i % 2])
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(nb):
        net(u8[i % 2])
    torch.cuda.synchronize()
    v_blk = world * batch * nb / dist.max_over_ranks(time.perf_counter() - t0)
    return {'value': v_u8, 'unit': UNIT, 'h2d_bytes_per_step': int(u8[0].nbytes), 'd2h_bytes_per_step': d2h,
            'steps': n_batches, 'input': 'uint8 NCHW images in pinned host memory (the first layer converts; the numpy '
                                         'reference computes on x.astype(float16) for such an input)',
            'api': 'for y in net.map(batches): pinned numpy batch in, numpy result out, 2 batches in flight',
            'result_views': None if v_views is None else {
                'value': v_views, 'unit': UNIT, 'steps': n_batches,
                'api': 'net.map(batches, copy=False): results are views of the pinned ring (valid for depth + 1 further results) '
                       'instead of fresh numpy arrays -- no %d MB host memcpy per batch' % (d2h // 1000000)},
            'fp16_host': {'value': v_f16, 'unit': UNIT, 'h2d_bytes_per_step': int(f16[0].nbytes), 'steps': n_batches,
                          'input': '%s NCHW host batches' % str(hosts[0].dtype)},
            'blocking_call': {'value': v_blk, 'unit': UNIT, 'steps': nb,
                              'api': 'y = net(x) per uint8 batch (upload in two halves, one synchronisation)'}}


if __name__ == '__main__':
    main()
    try:
        import torch.distributed as _d
        if _d.is_available() and _d.is_initialized():
            _d.destroy_process_group()
    except Exception:
        pass
